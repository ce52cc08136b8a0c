"""Row a14 / f3: the point-feature (RGB) branch.

Parity is UNPINNED for this row: the reference's torch port of the branch does not run and
TensorFlow is absent (oracle/rgb.py).  CPU: the oracle's restatement is tied to the PINNED
closed form wherever the two overlap (unit features reproduce the occupancy grid, per-channel
blur = the occupancy blur, white features integrate to the total ray probability) and frozen by
the fixture tests/golden/rgb.npz.  GPU: the kernels through the Python mirror -> ctypes -> C
ABI against the oracle (forward 1e-5, gradients 1e-4, scale-relative)."""
import os

import numpy as np
import pytest
import torch

import _golden
import _inputs
from golden.make_golden_rgb import CASES, make_inputs, run
from oracle import closed_form as CF
from oracle import rgb as ORGB
from oracle.config import default_cfg

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rgb.npz")))


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
def test_cuda_matches_oracle_fixture(name, plane_local):
    import pytorch_unsup_pc_b200 as dpc
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
