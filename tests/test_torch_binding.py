"""The optional C++ autograd binding (csrc/torch_binding.cpp -> lib/dpc_b200_torch.so) against the
ctypes Function: both call the same C ABI, so every output and gradient must be bit-identical
(the default path is order-free)."""
import pytest
import torch

import _inputs
from oracle import closed_form as CF
from oracle.config import default_cfg


def test_binding_loads_and_matches_the_library():
    from pytorch_unsup_pc_b200 import _lib, _torch_binding
    ext = _torch_binding.load()
    if ext is None:
        pytest.skip("lib/dpc_b200_torch.so not built (python __graft_entry__.py builds it)")
    assert ext.abi_version() == _lib.load().dpc_version()
    assert "project" in dir(ext)


@pytest.mark.gpu
@pytest.mark.parametrize("outputs", [True, False])
@pytest.mark.parametrize("replicated", [False, True])
def test_binding_equals_ctypes_function(outputs, replicated):
    import pytorch_unsup_pc_b200 as dpc
    from pytorch_unsup_pc_b200 import _lib, _torch_binding, ops
    if _torch_binding.load() is None:
        pytest.skip("C++ binding not built")
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    P, N, R = 8, 3000, 4
    case = _inputs.make_case(cfg, P, N, 606, translation=True, focal=True, scale=True, screened=False)
    taps = ops.host_taps(CF.smoothing_taps(cfg, 2.0))
    pts = case["points"][: P // R].contiguous() if replicated else case["points"]
    g = torch.Generator().manual_seed(607)
    sel = torch.stack([torch.randperm(N, generator=g)[:2000] for _ in range(P)]).int().to(dev) if replicated else None
    n = 2000 if replicated else N
    rep = (R, N, sel) if replicated else None
    Wp, Wd = (w.to(dev).squeeze(-1) for w in _inputs.loss_weights(P, 64))
    Wt = torch.rand(P, n, 3, generator=g).to(dev)
    res = []
    for fn in (ops.project, ops.ProjectFn.apply):
        leaves = [pts.to(dev).requires_grad_(), case["quat"].to(dev).requires_grad_(),
                  case["translation"].to(dev).requires_grad_(),
                  case["focal"].reshape(-1).to(dev).requires_grad_(),
                  case["scale"].reshape(-1).to(dev).requires_grad_()]
        params = ops.make_params(cfg, P, n, flip_y=True)
        out = fn(*leaves, params, taps, outputs, outputs, _lib.SCATTER_ATOMIC, True, rep)
        loss = (out[0] * Wp).sum() + 0.1 * (out[1] * Wd).sum() + 0.01 * (out[2] * Wt).sum()
        if outputs:
            loss = loss + 1e-3 * out[3].sum() + 1e-3 * (out[4] * out[4]).sum()
        grads = torch.autograd.grad(loss, leaves)
        res.append((out, grads))
    (o1, g1), (o2, g2) = res
    for a, b in zip(o1, o2):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)
