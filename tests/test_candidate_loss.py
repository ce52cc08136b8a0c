"""Row f1: candidate-selection projection loss.
CPU: the oracle restatement against the reference's golden vectors (and the
reference itself when /root/reference is mounted).  GPU: the fused kernels,
through the Python mirror -> ctypes -> C ABI, against the same vectors."""
import os

import numpy as np
import pytest
import torch

from golden.make_golden_loss import CASES, make_inputs
from oracle import loss as OL
from oracle import ref_loader as RL
from oracle.config import default_cfg

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                 "candidate_loss.npz")))


def _case(name):
    spec = CASES[name]
    masks = torch.from_numpy(GOLD[name + "/masks"])
    projs = torch.from_numpy(GOLD[name + "/projs"])
    w = torch.from_numpy(GOLD[name + "/weights"]) if name + "/weights" in GOLD else None
    return spec, masks, projs, w


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    spec, masks, projs, w = _case(name)
    p = projs.double().requires_grad_()
    total, min_loss = OL.add_proj_loss(masks, p, spec["C"], spec["scale"], w)
    (gp,) = torch.autograd.grad(total, p)
    assert total.item() == pytest.approx(float(GOLD[name + "/total"]), rel=1e-12)
    assert min_loss.tolist() == GOLD[name + "/min_loss"].tolist()
    assert np.abs(gp.numpy() - GOLD[name + "/g_projs"]).max() <= 1e-12


def test_golden_inputs_are_reproducible():
    for name, spec in CASES.items():
        masks, projs, w = make_inputs(spec)
        assert np.array_equal(masks.numpy(), GOLD[name + "/masks"])
        assert np.array_equal(projs.numpy(), GOLD[name + "/projs"])


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    cfg = RL.reference_cfg(pose_predict_num_candidates=4, pose_predictor_student=False)
    g = torch.Generator().manual_seed(5)
    masks = (torch.rand(3, 1, 128, 128, generator=g) > 0.5).float()
    projs = torch.rand(12, 64, 64, 1, generator=g, dtype=torch.float64)
    a = projs.clone().requires_grad_()
    b = projs.clone().requires_grad_()
    t_ref, m_ref = RL.ref_candidate_loss(cfg, masks.clone(), a, 1.0)
    t_or, m_or = OL.add_proj_loss(masks, b, 4, 1.0)
    assert t_ref.item() == t_or.item() and m_ref.tolist() == m_or.tolist()
    assert torch.equal(torch.autograd.grad(t_ref, a)[0], torch.autograd.grad(t_or, b)[0])


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("i", range(10))
def test_oracle_matches_live_reference_sweep(i):
    """The pin widened to a seeded sweep: 2-5 candidates, 1-6 views, pooled and un-pooled ground
    truth, per-sample weights, fp32 and fp64 predictions, exact ties between candidates (the
    argmin must take the FIRST minimum, as torch.argmin does).  Loss, argmin and gradient are
    bit-identical to the reference's own methods."""
    g = torch.Generator().manual_seed(7300 + i)
    pick = lambda xs: xs[int(torch.randint(len(xs), (1,), generator=g))]   # noqa: E731
    C, BV, V = pick([2, 3, 4, 5]), pick([1, 2, 3, 6]), pick([16, 32, 64])
    G = V * pick([1, 2, 4])
    weighted, dtype, scale = pick([False, True]), pick([torch.float32, torch.float64]), pick([1.0, 0.37, 160.0])
    cfg = RL.reference_cfg(pose_predict_num_candidates=C, pose_predictor_student=False,
                           variable_num_views=weighted, vox_size=V)
    masks = (torch.rand(BV, 1, G, G, generator=g) > 0.5).float()
    projs = torch.rand(BV * C, V, V, 1, generator=g, dtype=dtype)
    if i % 2:                                   # candidates 0 and C-1 of the first view tie exactly
        projs[C - 1] = projs[0]
    w = torch.rand(BV, generator=g).round() if weighted else None
    a, b = projs.clone().requires_grad_(), projs.clone().requires_grad_()
    t_ref, m_ref = RL.ref_candidate_loss(cfg, masks.clone(), a, scale, w)
    t_or, m_or = OL.add_proj_loss(masks, b, C, scale, w)
    assert t_ref.item() == t_or.item() and m_ref.tolist() == m_or.tolist()
    assert torch.equal(torch.autograd.grad(t_ref, a)[0], torch.autograd.grad(t_or, b)[0])


def test_unsupported_branches_raise():
    import pytorch_unsup_pc_b200 as dpc
    z = {"masks": torch.zeros(2, 1, 8, 8)}
    o = {"projs": torch.zeros(2, 8, 8, 1)}
    with pytest.raises(NotImplementedError):
        dpc.add_proj_loss(default_cfg(pose_predict_num_candidates=1), z, o, 1.0)
    with pytest.raises(NotImplementedError):
        dpc.add_proj_loss(default_cfg(pose_predict_num_candidates=2, pose_predictor_student=True), z, o, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.add_proj_loss(default_cfg(pose_predict_num_candidates=2), z, o, 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name):
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    spec, masks, projs, w = _case(name)
    cfg = default_cfg(pose_predict_num_candidates=spec["C"], variable_num_views=w is not None)
    p = projs.to(dev).requires_grad_()
    inputs = {"masks": masks.to(dev)}
    if w is not None:
        inputs["valid_samples"] = w.to(dev)
    total, min_loss = dpc.add_proj_loss(cfg, inputs, {"projs": p}, spec["scale"])
    (gp,) = torch.autograd.grad(total * 3.0, p)          # a non-unit upstream gradient
    ref = float(GOLD[name + "/total"])
    assert abs(total.item() - ref) / abs(ref) < 1e-5
    assert min_loss.cpu().tolist() == GOLD[name + "/min_loss"].tolist()
    g_ref = 3.0 * GOLD[name + "/g_projs"]
    err = np.abs(gp.cpu().numpy().astype(np.float64) - g_ref).max() / np.abs(g_ref).max()
    assert err < 1e-5, err
    # the un-fused entry point (already pooled gt, [BV,V,V,1]) agrees
    gt = OL.pool_gt(masks, spec["V"]).contiguous().to(dev)
    t2, m2 = dpc.proj_loss_pose_candidates(cfg, gt, p, inputs)
    assert abs(t2.item() * spec["scale"] - ref) / abs(ref) < 1e-5
    assert torch.equal(m2, min_loss)
