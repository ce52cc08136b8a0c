"""Row f2: replica-aware projection + point dropout on the device.
CPU: the oracle restatement against the reference's golden vectors (and the reference itself
when /root/reference is mounted).  GPU: the kernels through the Python mirror -> ctypes -> C
ABI against the same vectors, against the materialised-replica path, and the sampler's
properties."""
import os

import numpy as np
import pytest
import torch

import _golden
import _inputs
from golden.make_golden_replicas import CASES, make_inputs
from oracle import closed_form as CF
from oracle import ref_loader as RL
from oracle import replicas as OR
from oracle.config import default_cfg

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                 "replicas.npz")))


def _case(name, device="cpu"):
    spec = CASES[name]
    cfg = default_cfg(**spec["cfg"])
    t = {k: torch.from_numpy(GOLD["%s/in_%s" % (name, k)]).to(device) for k in ("points", "quat", "scale")}
    kx, ky, kz = (torch.from_numpy(GOLD["%s/taps_%s" % (name, a)]) for a in "xyz")
    kernel = [kx.reshape(1, 1, 1, 1, -1), ky.reshape(1, 1, 1, -1, 1), kz.reshape(1, 1, -1, 1, 1)]
    idx = torch.from_numpy(GOLD[name + "/indices"]).long().to(device) if name + "/indices" in GOLD else None
    return spec, cfg, t, kernel, idx


def _pairs(idx):
    """[P,M] point indices -> the reference sampler's [P,M,2] (row, index) layout."""
    rows = torch.arange(idx.shape[0]).reshape(-1, 1).expand_as(idx)
    return torch.stack([rows, idx], dim=-1)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    spec, cfg, t, kernel, idx = _case(name)
    R = spec["views"] * spec["cands"]
    leaves = [t[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
    out = OR.project_replicated(cfg, leaves[0], leaves[1], R, None if idx is None else _pairs(idx),
                                None, kernel, leaves[2])
    P = t["quat"].shape[0]
    Wp, Wd = _inputs.loss_weights(P, cfg.vox_size)
    loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
    grads = torch.autograd.grad(loss, leaves)
    for k in ("proj", "proj_depth", "tr_pc"):
        assert _golden.rel_err(out[k], GOLD[name + "/" + k]) < 1e-13, k
    assert abs(loss.item() - float(GOLD[name + "/loss"])) < 1e-9 * abs(float(GOLD[name + "/loss"]))
    for k, g in zip(("points", "quat", "scale"), grads):
        assert _golden.rel_err(g, GOLD[name + "/grad_" + k]) < 2e-6, k


def test_golden_inputs_are_reproducible():
    for name, spec in CASES.items():
        _, pts, quat, scale = make_inputs(spec)
        assert np.array_equal(pts.numpy(), GOLD[name + "/in_points"])
        assert np.array_equal(quat.numpy(), GOLD[name + "/in_quat"])


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    """tf_repeat_0, the sampler (same numpy stream -> same indices) and the gather, against
    the reference's own functions; then the whole replicated projection, bit for bit."""
    RL.load()
    import models.model_pc_to as model_pc
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 5, 2, generator=g)
    assert torch.equal(model_pc.tf_repeat_0(x, 4), OR.tf_repeat_0(x, 4))
    np.random.seed(11)
    mine = OR.sample_indices(6, 50, 0.37)
    pts = torch.rand(6, 50, 3, generator=g)
    np.random.seed(11)
    ref_pts, _ = RL.load()["pc_to"].pc_point_dropout(pts, None, 0.37)
    assert torch.equal(ref_pts, OR.pc_point_dropout(pts, None, mine)[0])
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    case = _inputs.make_case(cfg, 8, 300, 77, screened=False)
    cloud = case["points"][:2].contiguous()
    kernel = RL.ref_smoothing_kernel(cfg, 1.5)
    a = cloud.clone().requires_grad_()
    b = cloud.clone().requires_grad_()
    ref, idx = RL.ref_project_replicated(cfg, a, case["quat"], 2, 2, 0.8, 5, None, kernel, case["scale"])
    ora = OR.project_replicated(cfg, b, case["quat"], 4, _pairs(idx), None, kernel, case["scale"])
    for k in ("proj", "proj_depth", "tr_pc"):
        assert torch.equal(ref[k], ora[k]), k
    ga, = torch.autograd.grad(ref["proj"].sum() + ref["proj_depth"].sum(), a)
    gb, = torch.autograd.grad(ora["proj"].sum() + ora["proj_depth"].sum(), b)
    assert _golden.rel_err(gb, ga) < 2e-6


def test_host_validation():
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.pointcloud_project_replicated(cfg, torch.zeros(2, 10, 3), torch.zeros(4, 4), None, None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.pc_point_dropout(torch.zeros(2, 10, 3), None, 0.5)


# ---------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("plane_local", [True, False])
def test_cuda_matches_golden(name, plane_local):
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    spec, cfg, t, kernel, idx = _case(name, dev)
    leaves = [t[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
    with dpc.options(plane_local=plane_local):
        out = dpc.pointcloud_project_replicated(cfg, leaves[0], leaves[1], None, None, kernel,
                                                scaling_factor=leaves[2], indices=idx)
        P = t["quat"].shape[0]
        Wp, Wd = (w.to(dev) for w in _inputs.loss_weights(P, cfg.vox_size))
        loss = (out["proj"] * Wp).sum() + 0.1 * (out["proj_depth"] * Wd).sum()
        grads = torch.autograd.grad(loss, leaves)
    for k in ("proj", "proj_depth", "tr_pc"):
        assert _golden.rel_err(out[k], GOLD[name + "/" + k]) < 1e-5, k       # forward: 1e-5
    assert grads[0].shape == t["points"].shape                                 # the CLOUD's gradient
    for k, g in zip(("points", "quat", "scale"), grads):
        assert _golden.rel_err(g, GOLD[name + "/grad_" + k]) < 1e-4, k        # gradients: 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("deterministic", [False, True])
def test_replicated_equals_materialised(deterministic):
    """Same kernels, same per-projection inputs: the forward is bit-identical to
    pointcloud_project_fast on the materialised copies in the deterministic mode (and to
    scatter-order rounding otherwise); the cloud gradient is the replica-ordered sum of the
    per-copy gradients, routed through the selection."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    B, R, N, M = 3, 4, 3000, 2100
    g = torch.Generator().manual_seed(21)
    cloud = ((torch.rand(B, N, 3, generator=g) - 0.5) * 0.9).to(dev)
    quat = torch.randn(B * R, 4, generator=g).to(dev).requires_grad_()
    scale = (0.2 + 0.8 * torch.rand(B * R, 1, generator=g)).to(dev)
    kernel = dpc.smoothing_kernel(cfg, 2.0)
    Wp, Wd = (w.to(dev) for w in _inputs.loss_weights(B * R, 64))
    dpc.set_deterministic(deterministic)
    try:
        for sel in (None, torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(B * R)]).to(dev)):
            a = cloud.clone().requires_grad_()
            out = dpc.pointcloud_project_replicated(cfg, a, quat, None, None, kernel,
                                                    scaling_factor=scale, indices=sel)
            ga, gq = torch.autograd.grad((out["proj"] * Wp).sum() + 0.1 * (out["proj_depth"] * Wd).sum(),
                                         [a, quat])
            copies = OR.tf_repeat_0(cloud, R)
            if sel is not None:
                copies = torch.gather(copies, 1, sel.long().unsqueeze(-1).expand(-1, -1, 3))
            c = copies.contiguous().requires_grad_()
            ref = dpc.pointcloud_project_fast(cfg, c, quat, None, None, kernel, scaling_factor=scale)
            gc, gq2 = torch.autograd.grad((ref["proj"] * Wp).sum() + 0.1 * (ref["proj_depth"] * Wd).sum(),
                                          [c, quat])
            assert torch.equal(out["tr_pc"], ref["tr_pc"])
            if deterministic:
                for k in ("proj", "proj_depth", "voxels", "drc_probs"):
                    assert torch.equal(out[k], ref[k]), k
                assert torch.equal(gq, gq2)
                # replica-ordered sum of the per-copy gradients
                if sel is None:
                    want = gc.reshape(B, R, N, 3)
                    acc = torch.zeros(B, N, 3, device=dev)
                    for r in range(R):
                        acc = acc + want[:, r]
                else:
                    acc = torch.zeros(B, N, 3, device=dev)
                    for r in range(R):
                        rows = torch.arange(B, device=dev) * R + r
                        part = torch.zeros(B, N, 3, device=dev)
                        part.scatter_(1, sel[rows].long().unsqueeze(-1).expand(-1, -1, 3), gc[rows])
                        acc = acc + part
                assert torch.equal(ga, acc)
            else:
                for k in ("proj", "proj_depth"):
                    assert _golden.rel_err(out[k], ref[k]) < 1e-5, k
    finally:
        dpc.set_deterministic(False)


@pytest.mark.gpu
def test_dropout_sampler_properties():
    """Exactly M distinct in-range indices per projection, ascending; a pure function of the
    seed; different projections / seeds draw different subsets; inclusion frequency M/N."""
    from pytorch_unsup_pc_b200 import ops
    dev = torch.device("cuda:0")
    for (P, N, M) in ((64, 8000, 5600), (5, 257, 256), (3, 100, 100), (4, 70000, 1), (7, 1000, 333)):
        sel = ops.dropout_indices(P, N, M, 1234, dev)
        assert sel.shape == (P, M) and sel.dtype == torch.int32
        s = sel.long()
        assert int(s.min()) >= 0 and int(s.max()) < N
        if M > 1:
            assert bool((s[:, 1:] > s[:, :-1]).all())              # ascending => distinct
        assert torch.equal(sel, ops.dropout_indices(P, N, M, 1234, dev))
        if M < N and N > 300:
            assert not torch.equal(sel, ops.dropout_indices(P, N, M, 1235, dev))
            assert not torch.equal(sel[0], sel[1])
    # inclusion frequency of every point over many projections: binomial(P, M/N)
    P, N, M = 4096, 500, 150
    sel = ops.dropout_indices(P, N, M, 99, dev).long()
    freq = torch.bincount(sel.reshape(-1), minlength=N).double() / P
    sigma = (0.3 * 0.7 / P) ** 0.5
    assert float((freq - 0.3).abs().max()) < 6 * sigma
    # pairwise independence proxy: neighbours are not picked together more often than chance
    hit = torch.zeros(P, N, dtype=torch.bool, device=dev)
    hit.scatter_(1, sel, True)
    both = (hit[:, 1:] & hit[:, :-1]).double().mean(0)
    assert float((both - 0.09).abs().max()) < 6 * (0.09 * 0.91 / P) ** 0.5 + 2e-3


@pytest.mark.gpu
def test_pc_point_dropout_op():
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(8)
    pts = torch.rand(6, 500, 3, generator=g).to(dev).requires_grad_()
    rgb = torch.rand(6, 500, 3, generator=g).to(dev)
    # the reference's own [P,M,2] index layout
    np.random.seed(4)
    pairs = OR.sample_indices(6, 500, 0.6)
    out, out_rgb = dpc.pc_point_dropout(pts, rgb, 0.6, indices=pairs)
    want, want_rgb = OR.pc_point_dropout(pts.detach().cpu(), rgb.cpu(), pairs)
    assert torch.equal(out.cpu(), want) and torch.equal(out_rgb.cpu(), want_rgb)
    w = torch.rand(out.shape, generator=g).to(dev)
    (gp,) = torch.autograd.grad((out * w).sum(), pts)
    ref = torch.zeros(6, 500, 3)
    ref[pairs[:, :, 0], pairs[:, :, 1]] = w.cpu()
    assert torch.equal(gp.cpu(), ref)
    # device sampler: M = int(N * keep_prob) rows of the input, seed-reproducible
    a, _ = dpc.pc_point_dropout(pts, None, 0.6, seed=7)
    b, _ = dpc.pc_point_dropout(pts, None, 0.6, seed=7)
    assert a.shape == (6, 300, 3) and torch.equal(a, b)
    torch.manual_seed(5)
    c, _ = dpc.pc_point_dropout(pts, None, 0.6)
    torch.manual_seed(5)
    d, _ = dpc.pc_point_dropout(pts, None, 0.6)
    assert torch.equal(c, d) and not torch.equal(a, c)


@pytest.mark.gpu
def test_replicated_with_device_dropout_full_size():
    """Workload-A shapes (16 clouds x 4 candidates, 8000 points, keep 0.7): the sampled
    selection is returned, and feeding it back through the materialised path reproduces the
    outputs; dropped points get exactly zero gradient."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    g = torch.Generator().manual_seed(31)
    cloud = ((torch.rand(16, 8000, 3, generator=g) - 0.5) * 0.9).to(dev).requires_grad_()
    quat = torch.randn(64, 4, generator=g).to(dev)
    kernel = dpc.smoothing_kernel(cfg, 3.0)
    dpc.set_outputs(voxels=False, drc_probs=False)
    try:
        out = dpc.pointcloud_project_replicated(cfg, cloud, quat, None, None, kernel, keep_prob=0.7, seed=3)
        sel = out["dropout_indices"]
        assert sel.shape == (64, 5600)
        (gc,) = torch.autograd.grad(out["proj"].sum(), cloud)
        kept = torch.zeros(16, 8000, dtype=torch.bool, device=dev)
        for r in range(4):
            kept.scatter_(1, sel[r::4].long(), True)
        assert bool((gc[~kept] == 0).all()) and bool((gc[kept] != 0).any())
        copies = torch.gather(OR.tf_repeat_0(cloud.detach(), 4), 1, sel.long().unsqueeze(-1).expand(-1, -1, 3))
        ref = dpc.pointcloud_project_fast(cfg, copies.contiguous(), quat, None, None, kernel)
        assert torch.equal(out["tr_pc"], ref["tr_pc"])
        assert _golden.rel_err(out["proj"], ref["proj"]) < 1e-5
    finally:
        dpc.set_outputs(voxels=True, drc_probs=True)


@pytest.mark.gpu
def test_replicated_with_point_features():
    """Rows f2 and f3 together: un-replicated clouds AND un-replicated features, with dropout,
    against the oracle on the materialised copies (forward 1e-5, gradients 1e-4)."""
    import pytorch_unsup_pc_b200 as dpc
    from oracle import rgb as ORGB
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    B, R, N, M = 2, 3, 600, 450
    case = _inputs.make_case(cfg, B * R, N, 909, scale=True, screened=False)
    g = torch.Generator().manual_seed(910)
    cloud = case["points"][:B].contiguous()
    feat = torch.rand(B, N, 3, generator=g)
    sel = torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(B * R)])
    W = torch.rand(B * R, 32, 32, 3, generator=g)
    kern = CF.smoothing_taps(cfg, 1.5)
    for idx in (sel, None):
        a, fa = cloud.clone().requires_grad_(), feat.clone().requires_grad_()
        pts, col = OR.tf_repeat_0(a, R), OR.tf_repeat_0(fa, R)
        if idx is not None:
            pts, col = OR.pc_point_dropout(pts, col, _pairs(idx))
        ref = ORGB.project_rgb(cfg, pts, case["quat"], col, None, kern, case["scale"])
        gr = torch.autograd.grad((ref["proj_rgb"] * W.double()).sum() + ref["proj"].sum(), [a, fa])
        b, fb = cloud.to(dev).requires_grad_(), feat.to(dev).requires_grad_()
        out = dpc.pointcloud_project_replicated(cfg, b, case["quat"].to(dev), None, fb, kern,
                                                scaling_factor=case["scale"].to(dev),
                                                indices=None if idx is None else idx.to(dev))
        gc = torch.autograd.grad((out["proj_rgb"] * W.to(dev)).sum() + out["proj"].sum(), [b, fb])
        assert _golden.rel_err(out["proj_rgb"], ref["proj_rgb"]) < 1e-5
        assert _golden.rel_err(out["proj"], ref["proj"]) < 1e-5
        assert gc[0].shape == cloud.shape and gc[1].shape == feat.shape
        for x, y in zip(gc, gr):
            assert _golden.rel_err(x, y) < 1e-4
