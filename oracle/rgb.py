"""Oracle of SURVEY.md 8a row a14 / next-row f3: the point-feature (RGB) branch.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

PINNED against the reference's TensorFlow original.  The reference's torch port of this branch
does not run (util/point_cloud_to.py:64 ``updates_raw. tf.expand_dims`` AttributeError;
util/drc.py:137 ``torch.float63``) and TensorFlow is not installed, but the TF original
(util/point_cloud.py) is complete: ``oracle.ref_loader.ref_project_tf`` executes it UNMODIFIED
through ``oracle/tf_shim.py`` (a ``tensorflow`` namespace over torch, validated bit for bit against
the reference's torch port on the occupancy path), and this restatement agrees with it to fp64
rounding: forward 1e-13, fp64 gradients 1e-10 (tests/test_rgb.py, live and against the
reference-made fixture tests/golden/rgb_tf.npz).  It follows the TF original line by line:
  * feature scatter                 util/point_cloud.py:99-129   (interpolate_scatter3d)
  * clip before the blur, blur      util/point_cloud.py:244-249, 148-154 (convolve_rgb)
  * division by blurred occupancy   util/point_cloud.py:256-260
  * clip after the blur             util/point_cloud.py:261-262
  * Y flip + colour integral        util/point_cloud.py:275-277, util/drc.py:132-142
    (project_volume_rgb_integral: white background appended as the last ray event)
on top of ``closed_form`` (which IS pinned) for the pose, the occupancy grid and the DRC
probabilities.  Gradients: torch autograd.
"""
import torch

from . import closed_form as CF


def scatter_features(cfg, tr_pc, rgb, stop_points_gradient=False):
    """tr_pc [P,N,3], rgb [P,N,C] -> voxels_rgb raw [P,Vz,V,V,C] (fp64): every valid point adds
    w(corner) * rgb to its eight corners (point_cloud.py:99-121)."""
    vz, v = CF.grid_dims(cfg)
    P, N, _ = tr_pc.shape
    C = rgb.shape[-1]
    dims = torch.tensor([vz, v, v], dtype=torch.int64)
    inside = ((tr_pc >= -0.5) & (tr_pc <= 0.5)).all(dim=-1)
    g = (tr_pc + 0.5) * (dims - 1)
    base = torch.floor(g)
    frac = g - base
    if stop_points_gradient:                                          # :112-113
        frac = frac.detach()
    cell = base.long()[inside]
    frac = frac[inside]
    col = rgb.double()[inside]                                        # [M,C]
    b_idx = torch.arange(P).unsqueeze(1).expand(P, N)[inside]
    grid = torch.zeros(P, vz, v, v, C, dtype=torch.float64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                w = ((frac[:, 0] if dz else 1.0 - frac[:, 0])
                     * (frac[:, 1] if dy else 1.0 - frac[:, 1])
                     * (frac[:, 2] if dx else 1.0 - frac[:, 2]))
                iz, iy, ix = cell[:, 0] + dz, cell[:, 1] + dy, cell[:, 2] + dx
                ok = (iz < vz) & (iy < v) & (ix < v)                  # zero-weight +0.5 corners
                grid = grid.index_put((b_idx[ok], iz[ok], iy[ok], ix[ok]),
                                      w[ok].unsqueeze(-1) * col[ok], accumulate=True)
    return grid


def convolve_rgb(voxels_rgb, kernels):
    """[P,Vz,V,V,C]: the three 1-D blurs on every channel separately (point_cloud.py:148-154)."""
    chans = []
    for c in range(voxels_rgb.shape[-1]):
        chans.append(CF.blur3d(voxels_rgb[..., c].unsqueeze(1), kernels).squeeze(1))
    return torch.stack(chans, dim=-1)


def rgb_integral(probs, voxels_rgb):
    """probs [Vz+1,P,V,V,1], voxels_rgb [P,Vz,V,V,C] -> [P,V,V,C]: sum_k p_k c_k with a white
    background as the last event (drc.py:132-142)."""
    rgb = voxels_rgb.permute(1, 0, 2, 3, 4)                           # swap batch and z
    bg = torch.ones_like(rgb[:1])
    return (probs * torch.cat([rgb, bg], 0)).sum(0)


def project_rgb(cfg, points, quat, rgb, translation=None, kernels=None, scaling_factor=None,
                focal_length=None):
    """The whole projection with point features: ``closed_form.project`` + voxels_rgb
    [P,Vz,V,V,C] (Y-flipped) and proj_rgb [P,V,V,C] (point_cloud.py:229-290)."""
    out = CF.project(cfg, points, quat, translation, kernels, scaling_factor, focal_length,
                     drop_oob=True)
    vrgb = scatter_features(cfg, out["tr_pc"], rgb, cfg.pc_rgb_stop_points_gradient)
    if kernels is not None:                                           # :244-249
        if not cfg.pc_rgb_clip_after_conv:
            vrgb = torch.clamp(vrgb, 0.0, 1.0)
        vrgb = convolve_rgb(vrgb, kernels)
    if cfg.pc_rgb_divide_by_occupancies:                              # :256-260
        div = CF.blur3d(out["voxels_raw"].detach().unsqueeze(1), kernels).squeeze(1)
        vrgb = vrgb / (div.unsqueeze(-1) + cfg.pc_rgb_divide_by_occupancies_epsilon)
    if cfg.pc_rgb_clip_after_conv:                                    # :261-262
        vrgb = torch.clamp(vrgb, 0.0, 1.0)
    vrgb = torch.flip(vrgb, [2])                                      # :275
    out["voxels_rgb"] = vrgb
    out["proj_rgb"] = rgb_integral(out["drc_probs"], vrgb)            # :276
    return out
