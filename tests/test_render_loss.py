"""The renderer and its loss as one step (rows f2 -> path -> f1 fused;
``pytorch_unsup_pc_b200.project_candidates_loss`` over dpc_render_loss_fwd / _bwd).

CPU: the oracle composition against vectors made by the REAL reference's own composition
(tests/golden/make_golden_render_loss.py).  GPU: the fused op against the same vectors
(loss / projections 1e-5, gradients 1e-4, argmin equal), against the composition of the three
separate CUDA ops (bit-identical: same kernels, same winner arithmetic, the skipped candidates'
gradients are exact zeros), with the two half-chains of the winner-only backward (>= 64 views)
in the deterministic mode, and through the general saved-state path (raw grid in global memory).
"""
import os

import numpy as np
import pytest
import torch

import _golden
from golden.make_golden_render_loss import CASES
from oracle import render_loss as ORL
from oracle.config import default_cfg

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                 "render_loss.npz")))


def _case(name, device="cpu"):
    spec = CASES[name]
    cfg = default_cfg(pose_predict_num_candidates=spec["cands"], variable_num_views=spec["weights"],
                      **spec["cfg"])
    t = {k: torch.from_numpy(GOLD[name + "/in_" + k]).to(device) for k in ("points", "quat", "scale")}
    t["masks"] = torch.from_numpy(GOLD[name + "/in_masks"]).float().to(device)
    t["weights"] = (torch.from_numpy(GOLD[name + "/in_weights"]).to(device)
                    if name + "/in_weights" in GOLD else None)
    kx, ky, kz = (torch.from_numpy(GOLD[name + "/taps_" + k]) for k in "xyz")
    kernel = [kx.reshape(1, 1, 1, 1, -1), ky.reshape(1, 1, 1, -1, 1), kz.reshape(1, 1, -1, 1, 1)]
    idx = torch.from_numpy(GOLD[name + "/indices"]).long().to(device) if name + "/indices" in GOLD else None
    return spec, cfg, t, kernel, idx


def _pairs(idx):
    P, M = idx.shape
    return torch.stack([torch.arange(P).reshape(P, 1).expand(P, M), idx.cpu()], dim=-1)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    spec, cfg, t, kernel, idx = _case(name)
    leaves = [t[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
    loss, min_loss, proj = ORL.project_candidates_loss(
        cfg, leaves[0], leaves[1], t["masks"], spec["cands"], kernel, leaves[2],
        weight_scale=spec["wscale"], valid_samples=t["weights"],
        indices=None if idx is None else _pairs(idx))
    assert loss.item() == pytest.approx(float(GOLD[name + "/loss"]), rel=1e-12)
    assert min_loss.tolist() == GOLD[name + "/min_loss"].tolist()
    assert _golden.rel_err(proj.float(), GOLD[name + "/proj"]) < 1e-7
    for k, g in zip(("points", "quat", "scale"), torch.autograd.grad(loss, leaves)):
        assert _golden.rel_err(g, GOLD[name + "/grad_" + k]) < 2e-6, k


def test_host_validation():
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=32, pose_predict_num_candidates=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.project_candidates_loss(cfg, torch.zeros(2, 10, 3), torch.zeros(4, 4), None,
                                    torch.zeros(2, 1, 32, 32))
    with pytest.raises(NotImplementedError):
        dpc.project_candidates_loss(default_cfg(vox_size=32, pose_predict_num_candidates=1),
                                    torch.zeros(2, 10, 3), torch.zeros(4, 4), None,
                                    torch.zeros(2, 1, 32, 32))


# ---------------------------------------------------------------------------- GPU
def _fused(dpc, cfg, spec, t, kernel, idx):
    leaves = [t[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
    out = dpc.project_candidates_loss(cfg, leaves[0], leaves[1], None, t["masks"], kernel,
                                      scaling_factor=leaves[2], weight_scale=spec["wscale"],
                                      valid_samples=t["weights"], indices=idx)
    grads = torch.autograd.grad(out["loss"], leaves)
    return out, grads


def _composed(dpc, cfg, spec, t, kernel, idx):
    leaves = [t[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
    dpc.set_outputs(voxels=False, drc_probs=False)
    try:
        proj = dpc.pointcloud_project_replicated(cfg, leaves[0], leaves[1], None, None, kernel,
                                                 scaling_factor=leaves[2], indices=idx)["proj"]
    finally:
        dpc.set_outputs(voxels=True, drc_probs=True)
    inputs = {"masks": t["masks"]}
    if t["weights"] is not None:
        inputs["valid_samples"] = t["weights"]
    total, min_loss = dpc.add_proj_loss(cfg, inputs, {"projs": proj}, spec["wscale"])
    grads = torch.autograd.grad(total, leaves)
    return total, min_loss, proj, grads


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("variant", ["default", "deterministic", "global_grid", "deterministic_global_grid"])
def test_cuda_matches_golden(name, variant):
    """default / deterministic: the plane-local saved state (winner-only backward, dL/dmask built
    in the ray kernel), with the atomic and with the sort-then-segment plane build.  global_grid:
    the raw grid in global memory, whose saved state is the general layout -- the fused op then
    writes dL/dmask for all P projections and runs the general backward: the same results
    through the other path."""
    import pytorch_unsup_pc_b200 as dpc
    spec, cfg, t, kernel, idx = _case(name, torch.device("cuda:0"))
    with dpc.options(deterministic=variant.startswith("deterministic"),
                     plane_local=not variant.endswith("global_grid")):
        out, grads = _fused(dpc, cfg, spec, t, kernel, idx)
    assert abs(out["loss"].item() - float(GOLD[name + "/loss"])) <= 1e-5 * abs(float(GOLD[name + "/loss"]))
    assert out["min_loss"].tolist() == GOLD[name + "/min_loss"].tolist()
    assert _golden.rel_err(out["projs"], GOLD[name + "/proj"]) < 1e-5
    assert grads[0].shape == t["points"].shape
    for k, g in zip(("points", "quat", "scale"), grads):
        assert _golden.rel_err(g, GOLD[name + "/grad_" + k].reshape(g.shape)) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_fused_equals_composition_of_the_three_ops(name):
    import pytorch_unsup_pc_b200 as dpc
    spec, cfg, t, kernel, idx = _case(name, torch.device("cuda:0"))
    out, grads = _fused(dpc, cfg, spec, t, kernel, idx)
    total, min_loss, proj, cgrads = _composed(dpc, cfg, spec, t, kernel, idx)
    assert torch.equal(out["projs"], proj)
    assert torch.equal(out["min_loss"], min_loss)
    assert abs(out["loss"].item() - total.item()) <= 2e-6 * abs(total.item())   # summation order
    for k, a, b in zip(("points", "quat", "scale"), grads, cgrads):
        assert torch.equal(a, b), (k, (a - b).abs().max().item())
    # the losing candidates' pose gradients are exact zeros
    C = spec["cands"]
    lose = torch.ones(t["quat"].shape[0], dtype=torch.bool, device=proj.device)
    lose[torch.arange(min_loss.numel(), device=proj.device) * C + min_loss] = False
    assert float(grads[1][lose].abs().max()) == 0.0 and float(grads[2][lose].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("B,views,C,N,V,keep", [(16, 4, 2, 300, 32, 1.0), (32, 2, 4, 257, 32, 0.6),
                                                (16, 1, 4, 8000, 64, 1.0)])
def test_large_batches_two_half_chains(B, views, C, N, V, keep):
    """>= 64 projections: the forward runs as two half-batches; >= 64 views: the winner-only
    backward runs as two half-chains.  Workload-A shapes (16 clouds x 4 candidates, 8000 points,
    64^3) included.  Checked against the composition of the separate ops."""
    import pytorch_unsup_pc_b200 as dpc
    from oracle import closed_form as CF
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=21 if V == 64 else 11,
                      pose_predict_num_candidates=C)
    g = torch.Generator().manual_seed(B * 1000 + N)
    P, BV = B * views * C, B * views
    t = {"points": ((torch.rand(B, N, 3, generator=g) - 0.5) * 0.9).to(dev),
         "quat": torch.randn(P, 4, generator=g).to(dev),
         "scale": (0.2 + 0.8 * torch.rand(P, 1, generator=g)).to(dev),
         "masks": (torch.rand(BV, 1, 2 * V, 2 * V, generator=g) > 0.6).float().to(dev),
         "weights": None}
    kernel = CF.smoothing_taps(cfg, 3.0 if V == 64 else 1.5)
    spec = {"wscale": 1.0, "cands": C}
    idx = None
    if keep < 1:
        from pytorch_unsup_pc_b200 import ops
        idx = ops.dropout_indices(P, N, int(N * keep), 77, dev)
    out, grads = _fused(dpc, cfg, spec, t, kernel, idx)
    total, min_loss, proj, cgrads = _composed(dpc, cfg, spec, t, kernel, idx)
    assert torch.equal(out["projs"], proj) and torch.equal(out["min_loss"], min_loss)
    assert abs(out["loss"].item() - total.item()) <= 2e-6 * abs(total.item())
    for k, a, b in zip(("points", "quat", "scale"), grads, cgrads):
        assert torch.equal(a, b), (k, (a - b).abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("keep", [1.0, 0.5])
def test_all_optional_inputs_against_oracle_composition(keep):
    """Translation, per-projection focal length, occupancy scale, per-view weights, dropout and a
    non-unit weight_scale together, on screened inputs: the fused op against the oracle
    composition (oracle.replicas + closed_form + loss, fp64) -- loss / projections 1e-5,
    every gradient 1e-4, argmin equal."""
    import pytorch_unsup_pc_b200 as dpc
    import _inputs
    from oracle import closed_form as CF
    from oracle.replicas import tf_repeat_0
    dev = torch.device("cuda:0")
    B, views, C, N, V, G = 3, 2, 3, 900, 32, 64
    R, P, BV = views * C, B * views * C, B * views
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=11, pose_predict_num_candidates=C,
                      variable_num_views=True)
    case = _inputs.make_case(cfg, P, N, 3100 + int(10 * keep), translation=True, focal=True,
                             scale=True, screened=False)
    g = torch.Generator().manual_seed(91)
    pts = case["points"][:B].contiguous()
    for _ in range(50):         # screen every cloud against all of its replicas' poses
        tr = CF.pose_transform(cfg, tf_repeat_0(pts, R), case["quat"], case["translation"], case["focal"])
        bad = _inputs.near_boundary(cfg, tr).reshape(B, R, -1).any(dim=1)
        if int(bad.sum()) == 0:
            break
        pts[bad] = (torch.rand(int(bad.sum()), 3, generator=g) - 0.5) * 0.9
    masks = (torch.rand(BV, 1, G, G, generator=g) > 0.55).float()
    weights = (torch.rand(BV, generator=g) > 0.3).float() + 0.25
    kernel = CF.smoothing_taps(cfg, 1.2)
    idx = None
    if keep < 1:
        M = int(N * keep)
        idx = torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(P)])
    names = ("points", "quat", "translation", "focal", "scale")
    cpu = dict(points=pts, quat=case["quat"], translation=case["translation"], focal=case["focal"],
               scale=case["scale"])
    lo = {k: v.clone().requires_grad_() for k, v in cpu.items()}
    loss_o, min_o, proj_o = ORL.project_candidates_loss(
        cfg, lo["points"], lo["quat"], masks, C, kernel, lo["scale"], lo["translation"], lo["focal"],
        weight_scale=0.7, valid_samples=weights, indices=None if idx is None else _pairs(idx))
    g_o = torch.autograd.grad(loss_o, [lo[k] for k in names])
    lc = {k: v.to(dev).requires_grad_() for k, v in cpu.items()}
    out = dpc.project_candidates_loss(cfg, lc["points"], lc["quat"], lc["translation"], masks.to(dev),
                                      kernel, scaling_factor=lc["scale"], focal_length=lc["focal"],
                                      weight_scale=0.7, valid_samples=weights.to(dev),
                                      indices=None if idx is None else idx.to(dev))
    g_c = torch.autograd.grad(out["loss"], [lc[k] for k in names])
    assert out["min_loss"].tolist() == min_o.tolist()
    assert abs(out["loss"].item() - loss_o.item()) <= 1e-5 * abs(loss_o.item())
    assert _golden.rel_err(out["projs"], proj_o.float()) < 1e-5
    for k, a, b in zip(names, g_c, g_o):
        assert _golden.rel_err(a, b.reshape(a.shape)) < 1e-4, k


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(),
                    reason="reference tree not mounted")
def test_oracle_composition_matches_live_reference():
    """The oracle composition against the REAL reference's own functions executed here
    (tf_repeat_0 + pc_point_dropout + projection in its CUDA-branch order + add_proj_loss), on a
    case no fixture holds: bit-identical loss and argmin, gradients to fp32 rounding."""
    import _inputs
    from oracle import ref_loader as RL
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11, pose_predict_num_candidates=2)
    B, views, C, N = 2, 2, 2, 400
    case = _inputs.make_case(cfg, B * views * C, N, 4711, scale=True, screened=False)
    cloud = case["points"][:B].contiguous()
    g = torch.Generator().manual_seed(3)
    masks = (torch.rand(B * views, 1, 64, 64, generator=g) > 0.5).float()
    kernel = RL.ref_smoothing_kernel(cfg, 1.5)
    rcfg = RL.reference_cfg(pose_predict_num_candidates=C, pose_predictor_student=False,
                            vox_size=32, pc_gauss_kernel_size=11)
    a = [t.clone().requires_grad_() for t in (cloud, case["quat"], case["scale"])]
    out, idx = RL.ref_project_replicated(rcfg, a[0], a[1], views, C, 0.75, 11, None, kernel, a[2])
    t_ref, m_ref = RL.ref_candidate_loss(rcfg, masks.clone(), out["proj"], 1.0)
    g_ref = torch.autograd.grad(t_ref, a)
    b = [t.clone().requires_grad_() for t in (cloud, case["quat"], case["scale"])]
    t_or, m_or, _ = ORL.project_candidates_loss(cfg, b[0], b[1], masks, C, kernel, b[2],
                                                indices=_pairs(idx))
    g_or = torch.autograd.grad(t_or, b)
    assert t_ref.item() == t_or.item() and m_ref.tolist() == m_or.tolist()
    for x, y in zip(g_or, g_ref):
        assert _golden.rel_err(x, y) < 2e-6


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(),
                    reason="reference tree not mounted")
@pytest.mark.parametrize("i", range(8))
def test_oracle_composition_matches_live_reference_sweep(i):
    """The same pin over a seeded sweep of what the fused step takes: 1-3 clouds, 1-3 views, 2-4
    candidates, 32^3 / 64^3, pooled and un-pooled masks, with and without point dropout,
    translation, focal length, per-sample weights and a loss scale.  Loss and argmin bit-identical
    to the reference's own functions, gradients to the fp32 rounding of the leaves."""
    import _inputs
    from oracle import ref_loader as RL
    g = torch.Generator().manual_seed(8800 + i)
    pick = lambda xs: xs[int(torch.randint(len(xs), (1,), generator=g))]   # noqa: E731
    B, views, C, N, V = pick([1, 2, 3]), pick([1, 2, 3]), pick([2, 3, 4]), pick([60, 250, 500]), pick([32, 32, 64])
    K, sigma, keep = pick([5, 11, 21]), pick([0.5, 1.5, 3.0]), pick([1.0, 0.75, 0.3])
    G, scale_w, weighted = V * pick([1, 2]), pick([1.0, 160.0]), pick([False, True])
    tr, fo = pick([False, True]), pick([False, True])
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=K, pose_predict_num_candidates=C,
                      variable_num_views=weighted)
    P = B * views * C
    case = _inputs.make_case(cfg, P, N, 8900 + i, scale=True, translation=tr, focal=fo, screened=False)
    cloud = case["points"][:B].contiguous()
    masks = (torch.rand(B * views, 1, G, G, generator=g) > 0.5).float()
    w = torch.rand(B * views, generator=g).round() if weighted else None
    kernel = RL.ref_smoothing_kernel(cfg, sigma)
    rcfg = RL.reference_cfg(pose_predict_num_candidates=C, pose_predictor_student=False, vox_size=V,
                            pc_gauss_kernel_size=K, variable_num_views=weighted)
    names = ["points", "quat", "scale"] + (["translation"] if tr else []) + (["focal"] if fo else [])
    src = dict(case, points=cloud)
    res = []
    for which in ("reference", "oracle"):
        lv = {k: src[k].clone().requires_grad_() for k in names}
        if which == "reference":
            out, idx = RL.ref_project_replicated(rcfg, lv["points"], lv["quat"], views, C, keep, 40 + i,
                                                 lv.get("translation"), kernel, lv["scale"], lv.get("focal"))
            total, mins = RL.ref_candidate_loss(rcfg, masks.clone(), out["proj"], scale_w, w)
        else:
            total, mins, _ = ORL.project_candidates_loss(
                cfg, lv["points"], lv["quat"], masks, C, kernel, lv["scale"], lv.get("translation"),
                lv.get("focal"), scale_w, w, indices=None if idx is None else _pairs(idx))
        res.append((total, mins, torch.autograd.grad(total, [lv[k] for k in names], allow_unused=True)))
    (t_ref, m_ref, g_ref), (t_or, m_or, g_or) = res
    assert t_ref.item() == t_or.item() and m_ref.tolist() == m_or.tolist()
    for k, x, y in zip(names, g_or, g_ref):
        assert (x is None) == (y is None), k
        if y is not None and float(y.abs().max()) > 0:
            assert _golden.rel_err(x, y) < 5e-6, k
