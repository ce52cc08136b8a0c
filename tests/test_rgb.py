"""Row a14 / f3: the point-feature (RGB) branch.

Pinning.  The reference's torch port of the branch does not run (util/point_cloud_to.py:64,
util/drc.py:137) and TensorFlow is absent, but the reference's TensorFlow ORIGINAL
(util/point_cloud.py:63-154, 229-290) is complete: it is executed here UNMODIFIED through
``oracle/tf_shim.py`` (a ``tensorflow`` namespace over torch) on top of the reference's own torch
``util/drc.py`` / ``util/quaternion.py`` (``oracle.ref_loader.ref_project_tf``).
  * the shim is validated first: without features the TF file through it reproduces the
    reference's torch port bit for bit, outputs and gradients;
  * ``oracle/rgb.py`` then agrees with the TF source to fp64 rounding (forward 1e-13, fp64
    gradients 1e-10), live in the build container and against the committed reference-made
    fixture ``tests/golden/rgb_tf.npz`` everywhere else;
  * the older oracle-made fixture ``rgb.npz`` equals the reference-made one entry by entry.
The oracle is also tied to the pinned closed form where the two overlap (unit features reproduce
the occupancy grid, per-channel blur = the occupancy blur, white features integrate to the total
ray probability).  GPU: the kernels through the Python mirror -> ctypes -> C ABI against both
fixtures and the oracle (forward 1e-5, gradients 1e-4, scale-relative)."""
import os

import numpy as np
import pytest
import torch

import _golden
import _inputs
from golden import make_golden_rgb_tf as TFGOLD
from golden.make_golden_rgb import CASES, make_inputs, run
from oracle import ref_loader as RL
from oracle import closed_form as CF
from oracle import rgb as ORGB
from oracle.config import default_cfg

_GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = dict(np.load(os.path.join(_GOLDEN_DIR, "rgb.npz")))          # made by oracle/rgb.py
GOLD_TF = dict(np.load(os.path.join(_GOLDEN_DIR, "rgb_tf.npz")))    # made by the reference's TF source
FIXTURES = {"oracle_made": GOLD, "reference_made": GOLD_TF}
needs_reference = pytest.mark.skipif(not RL.available(), reason="reference tree not present")


def _oracle_run(spec, dtype=None):
    """oracle/rgb.py on a case of make_golden_rgb_tf (all optional inputs), leaves cast to dtype."""
    cfg, case, rgb, W = TFGOLD.make_inputs(spec)
    leaves = {k: case[k].clone() for k in TFGOLD.INPUT_KEYS if case.get(k) is not None}
    leaves["rgb"] = rgb.clone()
    if dtype is not None:
        leaves = {k: v.to(dtype) for k, v in leaves.items()}
    leaves = {k: v.requires_grad_() for k, v in leaves.items()}
    kern = None if spec["sigma"] is None else CF.smoothing_taps(cfg, spec["sigma"])
    out = ORGB.project_rgb(cfg, leaves["points"], leaves["quat"], leaves["rgb"], leaves.get("translation"),
                           kern, leaves["scale"], leaves.get("focal"))
    loss = TFGOLD.loss_of(out, W, spec["P"], cfg.vox_size)
    return out, loss, dict(zip(leaves, torch.autograd.grad(loss, list(leaves.values()))))


@needs_reference
def test_tf_shim_reproduces_the_torch_port():
    """Validates oracle/tf_shim.py: the reference's TF file run through it WITHOUT features must
    equal the reference's own torch port (ref_loader.ref_project) bit for bit -- outputs and
    gradients, with every optional input, with and without the blur."""
    for v, k, sigma, seed in ((32, 11, 1.5, 31), (64, 21, 3.0, 32), (32, 11, None, 33)):
        cfg = default_cfg(vox_size=v, pc_gauss_kernel_size=k)
        case = _inputs.make_case(cfg, 2, 400, seed, translation=True, focal=True, scale=True, screened=True)
        kern = None if sigma is None else RL.ref_smoothing_kernel(cfg, sigma)
        Wp, Wd = _inputs.loss_weights(2, v)
        res = []
        for fn in (RL.ref_project,
                   lambda c, p, q, t, kk, s, f: RL.ref_project_tf(c, p, q, t, None, kk, s, f)):
            lv = [case[n].clone().requires_grad_() for n in ("points", "quat", "translation", "scale", "focal")]
            out = fn(cfg, lv[0], lv[1], lv[2], kern, lv[3], lv[4])
            loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
            res.append((out, torch.autograd.grad(loss, lv)))
        (a, ga), (b, gb) = res
        for key in ("proj", "proj_depth", "voxels", "tr_pc", "drc_probs"):
            assert torch.equal(a[key], b[key]), key
        assert b["voxels_rgb"] is None and b["proj_rgb"] is None
        for x, y in zip(ga, gb):
            assert torch.equal(x, y)


@needs_reference
@pytest.mark.parametrize("name", sorted(TFGOLD.CASES))
def test_oracle_matches_live_reference_tf_source(name):
    """oracle/rgb.py against the reference's TF source executed live: fp32 leaves as the fixture
    holds them (forward 1e-13; gradients to the fp32 rounding of the leaves' gradient casts), and
    fp64 leaves (gradients 1e-10)."""
    spec = TFGOLD.CASES[name]
    out, loss, grads = _oracle_run(spec)
    ref, rloss, rgrads = TFGOLD.run_reference(spec)
    assert ref["voxels_rgb"].shape == out["voxels_rgb"].shape
    for key in ("proj_rgb", "voxels_rgb", "proj", "proj_depth", "tr_pc"):
        assert _golden.rel_err(out[key].detach(), ref[key].detach()) < 1e-13, key
    assert abs(loss.item() - rloss.item()) <= 1e-12 * abs(rloss.item())
    for k in grads:
        assert _golden.rel_err(grads[k], rgrads[k]) < 5e-6, k
    _, _, g64 = _oracle_run(spec, torch.float64)
    _, _, r64 = TFGOLD.run_reference(spec, torch.float64)
    for k in g64:
        assert _golden.rel_err(g64[k], r64[k]) < 1e-10, k


@pytest.mark.parametrize("name", sorted(TFGOLD.CASES))
def test_oracle_matches_reference_made_fixture(name):
    """The same comparison against the committed reference-made vectors (runs anywhere)."""
    out, loss, grads = _oracle_run(TFGOLD.CASES[name])
    g = GOLD_TF
    assert abs(loss.item() - float(g[name + "/loss"])) <= 1e-12 * abs(float(g[name + "/loss"]))
    assert _golden.rel_err(out["proj_rgb"].detach(), g[name + "/proj_rgb"]) < 1e-13
    flat = out["voxels_rgb"].detach().reshape(-1)
    assert abs(flat.sum().item() - float(g[name + "/voxels_rgb_sum"])) <= 1e-12 * abs(float(g[name + "/voxels_rgb_sum"]))
    assert _golden.rel_err(flat[::61].float(), g[name + "/voxels_rgb_sub"]) < 1e-7
    for k, x in grads.items():
        assert _golden.rel_err(x, g[name + "/grad_" + k]) < 5e-6, k


def test_oracle_made_fixture_equals_reference_made_fixture():
    """rgb.npz (what the GPU tests were first written against) entry by entry against rgb_tf.npz:
    forward values to fp64 rounding, gradients to the rounding of their fp32 casts."""
    assert set(GOLD) <= set(GOLD_TF)
    for key, a in GOLD.items():
        tol = 5e-6 if "/grad_" in key else 1e-13
        assert _golden.rel_err(torch.from_numpy(np.asarray(a, dtype=np.float64)),
                               np.asarray(GOLD_TF[key], dtype=np.float64)) <= tol, key


def test_oracle_agrees_with_the_pinned_closed_form():
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    case = _inputs.make_case(cfg, 2, 300, 91, screened=False)
    tr = CF.pose_transform(cfg, case["points"], case["quat"])
    ones = torch.ones(2, 300, 2)
    grid = ORGB.scatter_features(cfg, tr, ones)
    raw = CF.scatter_trilinear(cfg, tr, drop_oob=True)
    assert torch.equal(grid[..., 0], raw) and torch.equal(grid[..., 1], raw)
    kern = CF.smoothing_taps(cfg, 1.5)
    g = torch.Generator().manual_seed(3)
    v = torch.rand(2, 32, 32, 32, 3, generator=g, dtype=torch.float64)
    blurred = ORGB.convolve_rgb(v, kern)
    for c in range(3):
        assert torch.equal(blurred[..., c], CF.blur3d(v[..., c].unsqueeze(1), kern).squeeze(1))
    # white features: proj_rgb = sum of all ray-event probabilities (e^c factors on the ends)
    out = CF.project(cfg, case["points"], case["quat"], None, kern, case["scale"])
    white = torch.ones(2, 32, 32, 32, 3, dtype=torch.float64)
    total = out["drc_probs"].sum(0)
    assert torch.allclose(ORGB.rgb_integral(out["drc_probs"], white), total.expand(-1, -1, -1, 3),
                          rtol=0, atol=1e-15)


def test_oracle_known_answer_single_point():
    """One point in the middle of a cell: the eight corners get rgb / 8; with no blur and full
    occupancy scaling the pixel under it shows p-weighted colour over a white background."""
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    V = 32
    u = (torch.tensor([[[10.5, 20.5, 5.5]]]) / (V - 1) - 0.5).double()      # (z, y, x) grid 10.5, 20.5, 5.5
    rgb = torch.tensor([[[0.8, 0.4, 0.2]]])
    grid = ORGB.scatter_features(cfg, u, rgb)
    assert grid.shape == (1, V, V, V, 3)
    corner = grid[0, 10:12, 20:22, 5:7]
    assert torch.allclose(corner, (rgb.double() / 8).expand(2, 2, 2, 3), atol=1e-12)
    assert abs(grid.sum().item() - rgb.double().sum().item()) < 1e-12


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_fixture(name):
    out, loss, grads = run(CASES[name])
    assert abs(loss.item() - float(GOLD[name + "/loss"])) <= 1e-9 * abs(float(GOLD[name + "/loss"]))
    assert _golden.rel_err(out["proj_rgb"], GOLD[name + "/proj_rgb"]) < 1e-12
    for k, g in grads.items():
        assert _golden.rel_err(g, GOLD[name + "/grad_" + k]) < 2e-6, k


def test_host_validation():
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.pointcloud_project_fast(cfg, torch.zeros(1, 8, 3), torch.ones(1, 4), None, torch.zeros(1, 8, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.convolve_rgb(cfg, torch.zeros(1, 32, 32, 32, 3), dpc.smoothing_kernel(cfg, 1.0))


# ---------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("plane_local", [True, False])
@pytest.mark.parametrize("fixture", sorted(FIXTURES))
def test_cuda_matches_fixture(name, plane_local, fixture):
    import pytorch_unsup_pc_b200 as dpc
    GOLD = FIXTURES[fixture]
    dev = torch.device("cuda:0")
    spec = CASES[name]
    cfg, case, rgb, W, kern = make_inputs(spec)
    leaves = {k: case[k].to(dev).requires_grad_() for k in ("points", "quat", "scale")}
    leaves["rgb"] = rgb.to(dev).requires_grad_()
    with dpc.options(plane_local=plane_local):
        out = dpc.pointcloud_project_fast(cfg, leaves["points"], leaves["quat"], None, leaves["rgb"],
                                          kern, scaling_factor=leaves["scale"])
        Wp, Wd = (w.to(dev) for w in _inputs.loss_weights(spec["P"], cfg.vox_size))
        loss = ((out["proj_rgb"] * W.to(dev)).sum() + (out["proj"] * Wp).sum()
                + 0.1 * (out["proj_depth"] * Wd).sum())
        grads = dict(zip(leaves, torch.autograd.grad(loss, list(leaves.values()))))
    assert out["proj_rgb"].shape == (spec["P"], cfg.vox_size, cfg.vox_size, 3)
    assert _golden.rel_err(out["proj_rgb"], GOLD[name + "/proj_rgb"]) < 1e-5            # forward: 1e-5
    flat = out["voxels_rgb"].detach().reshape(-1)
    assert abs(flat.double().sum().item() - float(GOLD[name + "/voxels_rgb_sum"])) < 1e-5 * abs(
        float(GOLD[name + "/voxels_rgb_sum"]))
    assert _golden.rel_err(flat[::61], GOLD[name + "/voxels_rgb_sub"]) < 1e-5
    for k, g in grads.items():
        assert _golden.rel_err(g, GOLD[name + "/grad_" + k]) < 1e-4, k                   # gradients: 1e-4


@pytest.mark.gpu
def test_cuda_matches_oracle_fresh_seed_chair_size():
    """64^3, K=21, 2000 points, features with 4 channels, all optional inputs."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    case = _inputs.make_case(cfg, 2, 2000, 777, translation=True, focal=True, scale=True, screened=True)
    g = torch.Generator().manual_seed(778)
    feat = torch.rand(2, 2000, 4, generator=g)
    W = torch.rand(2, 64, 64, 4, generator=g)
    kern = CF.smoothing_taps(cfg, 2.0)
    keys = ("points", "quat", "translation", "focal", "scale")
    a = {k: case[k].clone().requires_grad_() for k in keys}
    fa = feat.clone().requires_grad_()
    ref = ORGB.project_rgb(cfg, a["points"], a["quat"], fa, a["translation"], kern, a["scale"], a["focal"])
    gr = torch.autograd.grad((ref["proj_rgb"] * W.double()).sum(), list(a.values()) + [fa])
    b = {k: case[k].to(dev).requires_grad_() for k in keys}
    fb = feat.to(dev).requires_grad_()
    out = dpc.pointcloud_project_fast(cfg, b["points"], b["quat"], b["translation"], fb, kern,
                                      scaling_factor=b["scale"], focal_length=b["focal"])
    gc = torch.autograd.grad((out["proj_rgb"] * W.to(dev)).sum(), list(b.values()) + [fb])
    assert _golden.rel_err(out["proj_rgb"], ref["proj_rgb"]) < 1e-5
    assert _golden.rel_err(out["voxels_rgb"], ref["voxels_rgb"]) < 1e-5
    for k, x, y in zip(keys + ("rgb",), gc, gr):
        assert _golden.rel_err(x, y) < 1e-4, k


@pytest.mark.gpu
def test_standalone_mirrors():
    """pointcloud2voxels3d_fast with rgb, convolve_rgb, project_volume_rgb_integral."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    case = _inputs.make_case(cfg, 2, 300, 55, screened=True)
    tr = CF.pose_transform(cfg, case["points"], case["quat"]).float()
    g = torch.Generator().manual_seed(56)
    rgb = torch.rand(2, 300, 3, generator=g)
    vox, vrgb = dpc.pointcloud2voxels3d_fast(cfg, tr.to(dev), rgb.to(dev))
    want = ORGB.scatter_features(cfg, tr.double(), rgb)
    assert vrgb.shape == (2, 32, 32, 32, 3) and _golden.rel_err(vrgb, want) < 1e-5
    assert _golden.rel_err(vox, CF.scatter_trilinear(cfg, tr.double(), drop_oob=True)) < 1e-5
    kern = CF.smoothing_taps(cfg, 1.5)
    v = torch.rand(2, 32, 32, 32, 3, generator=g)
    vd = v.to(dev).requires_grad_()
    out = dpc.convolve_rgb(cfg, vd, kern)
    vc = v.double().requires_grad_()
    ref = ORGB.convolve_rgb(vc, kern)
    assert _golden.rel_err(out, ref) < 1e-5
    Wt = torch.rand(out.shape, generator=g)
    (ga,) = torch.autograd.grad((out * Wt.to(dev)).sum(), vd)
    (gb,) = torch.autograd.grad((ref * Wt.double()).sum(), vc)
    assert _golden.rel_err(ga, gb) < 1e-4
    p = torch.rand(33, 2, 32, 32, 1, generator=g)
    pd, cd = p.to(dev).requires_grad_(), v.to(dev).requires_grad_()
    pr = dpc.project_volume_rgb_integral(cfg, pd, cd)
    pc_, cc = p.double().requires_grad_(), v.double().requires_grad_()
    rr = ORGB.rgb_integral(pc_, cc)
    assert _golden.rel_err(pr, rr) < 1e-5
    Wi = torch.rand(pr.shape, generator=g)
    g1 = torch.autograd.grad((pr * Wi.to(dev)).sum(), [pd, cd])
    g2 = torch.autograd.grad((rr * Wi.double()).sum(), [pc_, cc])
    for x, y in zip(g1, g2):
        assert _golden.rel_err(x, y) < 1e-4
