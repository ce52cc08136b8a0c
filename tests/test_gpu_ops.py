"""GPU parity of the stand-alone ops (the reference's sub-functions) vs the oracle."""
import pytest
import torch

import _golden
import _inputs
from oracle import closed_form as CF
from oracle.config import default_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def dpc():
    import pytorch_unsup_pc_b200 as m
    m._lib.load()
    return m


def test_pc_perspective_transform(dpc):
    cfg = default_cfg(vox_size=64)
    case = _inputs.make_case(cfg, 5, 1000, 31, translation=True, focal=True, screened=False)
    keys = ("points", "quat", "translation", "focal")
    lo = {k: case[k].clone().requires_grad_() for k in keys}
    o = CF.pose_transform(cfg, lo["points"], lo["quat"], lo["translation"], lo["focal"])
    W = torch.rand(o.shape, generator=torch.Generator().manual_seed(1))
    go = torch.autograd.grad((o * W).sum(), list(lo.values()))
    lc = {k: case[k].to(DEV).requires_grad_() for k in keys}
    c = dpc.pc_perspective_transform(cfg, lc["points"], lc["quat"], lc["translation"], lc["focal"])
    gc = torch.autograd.grad((c * W.to(DEV)).sum(), list(lc.values()))
    # tr_pc is rounded from the same fp64 value the reference holds
    assert torch.equal(c.cpu(), o.float())
    for k, a, b in zip(keys, gc, go):
        assert _golden.rel_err(a, b.reshape(a.shape)) < 1e-4, k


@pytest.mark.parametrize("vox_size,vox_size_z", [(32, -1), (64, -1), (32, 16), (128, -1)])
def test_pointcloud2voxels3d_fast(dpc, vox_size, vox_size_z):
    cfg = default_cfg(vox_size=vox_size, vox_size_z=vox_size_z)
    g = torch.Generator().manual_seed(41)
    pc = ((torch.rand(3, 1200, 3, generator=g) - 0.5) * 1.1)
    vz = vox_size if vox_size_z == -1 else vox_size_z
    dims = torch.tensor([vz, vox_size, vox_size]).double()
    for _ in range(20):                                     # screen cell faces for the gradient
        gcoord = (pc.double() + 0.5) * (dims - 1)
        bad = ((gcoord - gcoord.round()).abs() < 1e-3).any(-1)
        if not bad.any():
            break
        pc[bad] = (torch.rand(int(bad.sum()), 3, generator=g) - 0.5) * 1.1
    assert not bad.any()
    po = pc.clone().requires_grad_()
    o = CF.scatter_trilinear(cfg, po.double())
    W = torch.rand(o.shape, generator=g).double()
    (go,) = torch.autograd.grad((o * o * W).sum() / 2, [po])
    for det in (False, True):
        pcg = pc.to(DEV).requires_grad_()
        with dpc.options(deterministic=det):
            c, rgb = dpc.pointcloud2voxels3d_fast(cfg, pcg, None)
        assert rgb is None
        (gc,) = torch.autograd.grad((c * c * W.float().to(DEV)).sum() / 2, [pcg])
        assert _golden.rel_err(c, o) < 1e-5, det
        assert _golden.rel_err(gc, go) < 1e-4, det


@pytest.mark.parametrize("vox_size,ksize,sigma", [(32, 11, 1.5), (64, 21, 3.0), (64, 21, 0.2),
                                                  (64, 5, 0.8), (128, 21, 2.0)])
def test_smoothen_voxels3d(dpc, vox_size, ksize, sigma):
    cfg = default_cfg(vox_size=vox_size, pc_gauss_kernel_size=ksize)
    g = torch.Generator().manual_seed(51)
    P = 2 if vox_size < 128 else 1
    vox = torch.rand(P, 1, vox_size, vox_size, vox_size, generator=g)
    kern = CF.smoothing_taps(cfg, sigma)
    vo = vox.clone().double().requires_grad_()
    o = CF.blur3d(vo, kern)
    W = torch.rand(o.shape, generator=g).double()
    (go,) = torch.autograd.grad((o * W).sum(), [vo])
    vc = vox.to(DEV).requires_grad_()
    c = dpc.smoothen_voxels3d(cfg, vc, kern)
    (gc,) = torch.autograd.grad((c * W.float().to(DEV)).sum(), [vc])
    assert c.shape == vox.shape
    assert _golden.rel_err(c, o) < 1e-5
    assert _golden.rel_err(gc, go) < 1e-4


def test_smoothen_voxels3d_asymmetric_taps(dpc):
    """The adjoint must use the reversed taps (not just the same ones)."""
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=5)
    g = torch.Generator().manual_seed(52)
    k = torch.tensor([0.1, 0.5, 0.2, 0.15, 0.05])
    kern = [k.reshape(1, 1, 1, 1, 5), (k * 1.1).reshape(1, 1, 1, 5, 1), (k * 0.9).reshape(1, 1, 5, 1, 1)]
    vox = torch.rand(1, 1, 32, 32, 32, generator=g)
    vo = vox.clone().double().requires_grad_()
    o = CF.blur3d(vo, kern)
    W = torch.rand(o.shape, generator=g).double()
    (go,) = torch.autograd.grad((o * W).sum(), [vo])
    vc = vox.to(DEV).requires_grad_()
    c = dpc.smoothen_voxels3d(cfg, vc, kern)
    (gc,) = torch.autograd.grad((c * W.float().to(DEV)).sum(), [vc])
    assert _golden.rel_err(c, o) < 1e-5
    assert _golden.rel_err(gc, go) < 1e-4


@pytest.mark.parametrize("logsum", [True, False])
def test_drc_projection_and_depth(dpc, logsum):
    cfg = default_cfg(vox_size=32, drc_logsum=logsum)
    g = torch.Generator().manual_seed(61)
    vox = torch.rand(2, 32, 32, 32, 1, generator=g) ** 4     # mostly empty, some dense
    vox[0, :, :4] = 0.0                                       # below the clip
    vox[1, 5, 10:14] = 1.0                                    # above 1 - clip
    vo = vox.clone().double().requires_grad_()
    po = CF.drc_probabilities(vo, cfg)
    mo, do = CF.drc_mask(po), CF.drc_depth(po, cfg)
    Wm, Wd = (torch.rand(mo.shape, generator=g).double() for _ in range(2))
    Wp = torch.rand(po.shape, generator=g).double()
    (go,) = torch.autograd.grad((mo * Wm).sum() + 0.1 * (do * Wd).sum() + (po * Wp).sum(), [vo])
    vc = vox.to(DEV).requires_grad_()
    mc, pc = dpc.drc_projection(vc, cfg)
    dc = dpc.drc_depth_projection(pc, cfg)
    (gc,) = torch.autograd.grad((mc * Wm.float().to(DEV)).sum() + 0.1 * (dc * Wd.float().to(DEV)).sum()
                                + (pc * Wp.float().to(DEV)).sum(), [vc])
    assert _golden.rel_err(mc, mo) < 1e-5
    assert _golden.rel_err(pc, po) < 1e-5
    assert _golden.rel_err(dc, do) < 1e-5
    assert _golden.rel_err(dpc.drc_event_probabilities(vc, cfg), po) < 1e-5
    assert _golden.rel_err(gc, go) < 1e-4


def test_no_cpu_fallback(dpc):
    cfg = default_cfg(vox_size=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.pointcloud_project_fast(cfg, torch.zeros(1, 4, 3), torch.ones(1, 4), None, None)


def test_plain_launches_give_the_same_bits(dpc):
    """dpc_set_programmatic_launch(0) launches the chain's kernels as plain stream-ordered kernels
    instead of programmatic dependent launches: one projection step (66 projections: the two
    half-batch chains on the internal streams) both ways, and every output and gradient must be
    the same bits."""
    lib = dpc._lib.load()
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    g = torch.Generator().manual_seed(77)
    pts = ((torch.rand(66, 3000, 3, generator=g) - 0.5) * 0.9).to(DEV)
    quat = torch.randn(66, 4, generator=g).to(DEV)
    kern = dpc.smoothing_kernel(cfg, 3.0)
    outs = []
    for pdl in (1, 0):
        lib.dpc_set_programmatic_launch(pdl)
        try:
            p, q = pts.clone().requires_grad_(), quat.clone().requires_grad_()
            o = dpc.pointcloud_project_fast(cfg, p, q, None, None, kern)
            gp, gq = torch.autograd.grad(o["proj"].sum() + o["proj_depth"].sum(), [p, q])
            torch.cuda.synchronize()
        finally:
            lib.dpc_set_programmatic_launch(-1)
        outs.append((o["proj"], o["proj_depth"], o["voxels"], o["drc_probs"], o["tr_pc"], gp, gq))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_full_size_properties(dpc):
    """Microbench shapes (P=64, N=8000, 64^3, K=21): size-independent properties."""
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    g = torch.Generator().manual_seed(1001)
    P, N = 64, 8000
    pts = ((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9).to(DEV).requires_grad_()
    quat = torch.randn(P, 4, generator=g).to(DEV).requires_grad_()
    scale = (0.2 + 0.8 * torch.rand(P, 1, generator=g)).to(DEV).requires_grad_()
    kern = CF.smoothing_taps(cfg, 3.0)
    out = dpc.pointcloud_project_fast(cfg, pts, quat, None, None, kern, scaling_factor=scale)
    probs = out["drc_probs"]
    # termination probabilities of a ray sum to 1 (up to the exp(c) end factors)
    total = probs.sum(0)
    assert (total - 1).abs().max().item() < 3e-5
    assert out["proj"].min().item() >= 0 and out["proj"].max().item() <= 1 + 1e-5
    assert torch.allclose(out["proj"], probs[:-1].sum(0), atol=2e-6)
    # depth lies between the near plane and max_depth
    assert out["proj_depth"].min().item() >= cfg.camera_distance - 0.5 - 1e-4
    assert out["proj_depth"].max().item() <= cfg.max_depth * (1 + 2e-5)
    # mass conservation of the scatter: one unit per in-frustum point
    tr = out["tr_pc"].detach()
    inside = ((tr >= -0.5) & (tr <= 0.5)).all(-1)
    raw, _ = dpc.pointcloud2voxels3d_fast(cfg, tr, None)
    assert torch.allclose(raw.sum((1, 2, 3)), inside.sum(1).float(), rtol=1e-5)
    # gradients are finite, zero for out-of-frustum points when only proj is used
    gp, gq, gs = torch.autograd.grad(out["proj"].sum(), [pts, quat, scale])
    assert torch.isfinite(gp).all() and torch.isfinite(gq).all() and torch.isfinite(gs).all()
    assert torch.count_nonzero(gp[~inside]).item() == 0
    assert torch.count_nonzero(gp[inside]).item() > 0.9 * inside.sum().item()
    # a quaternion and its positive multiple give the same projection
    out2 = dpc.pointcloud_project_fast(cfg, pts, quat * 3.0, None, None, kern, scaling_factor=scale)
    assert _golden.rel_err(out2["proj"], out["proj"]) < 1e-5
    # and dL/dq is orthogonal to q (scale invariance)
    assert ((gq * quat).sum(-1).abs() / (gq.norm(dim=-1) * quat.norm(dim=-1) + 1e-12)).max() < 1e-4
