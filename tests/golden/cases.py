"""The golden-vector cases: one row per committed ``<name>.npz``.

Shared by ``make_golden.py`` (which runs the REAL reference to produce the
files) and the tests (which replay the same seeded inputs through the oracle
and the CUDA path).  ``cfg`` entries override oracle.config.default_cfg().
"""

CASES = {
    # small grid, every optional input on (translation, focal tensor, scale)
    "small_v32_all_inputs": dict(
        cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=1.5, P=3, N=600, seed=1001,
        translation=True, focal=True, scale=True),
    # config-1 / microbench shapes (chair_unsupervised), two projections
    "chair_v64_sigma3": dict(
        cfg=dict(vox_size=64, pc_gauss_kernel_size=21), sigma=3.0, P=2, N=8000, seed=1002,
        scale=True),
    # end of the sigma schedule: 3 significant taps
    "chair_v64_sigma02": dict(
        cfg=dict(vox_size=64, pc_gauss_kernel_size=21), sigma=0.2, P=2, N=4000, seed=1003,
        scale=True),
    # clustered cloud: raw occupancy > 1 exercises the pre-blur clamp mask
    "clustered_v64": dict(
        cfg=dict(vox_size=64, pc_gauss_kernel_size=21), sigma=1.0, P=2, N=4000, seed=1004,
        kind="clustered", scale=True),
    # anisotropic Z grid (vox_size_z != -1) with separate, shorter Z taps.  The
    # reference's own smoothing_kernel crashes on this branch (gauss_kernel.py:49
    # reshapes the fsz_z taps to fsz), so the generator builds the intended
    # kernel list (:38-47) from the reference's gauss_kernel_1d by hand.
    "aniso_z16_v32": dict(
        cfg=dict(vox_size=32, vox_size_z=16, pc_gauss_kernel_size=11), sigma=1.5, P=2, N=1500,
        seed=1005, scale=True, aniso_kernel=True),
    # no blur (kernel=None, TF semantics) and no occupancy scaling
    "noblur_noscale_v32": dict(
        cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=None, P=2, N=1000, seed=1006,
        scale=False),
    # product-form DRC (drc_logsum: false)
    "drc_product_v32": dict(
        cfg=dict(vox_size=32, pc_gauss_kernel_size=11, drc_logsum=False), sigma=1.5, P=2, N=1000,
        seed=1007, scale=True),
    # paper scale grid, one projection
    "paper_v128": dict(
        cfg=dict(vox_size=128, pc_gauss_kernel_size=21), sigma=3.0, P=1, N=16000, seed=1008,
        scale=True),
}

# reference recipe run/pc_full_proj_test.py:48-61 (numpy seed 0, uniform [0,1))
RECIPE = dict(cfg=dict(vox_size=64, pc_gauss_kernel_size=21), sigma=3.0, P=128, N=140)

# subsampling stride for the big tensors kept in the fixtures
VOX_STRIDE = 61
