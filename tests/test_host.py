"""CPU: host logic and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from oracle.config import default_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import pytorch_unsup_pc_b200 as dpc
    header = open(os.path.join(ROOT, "include", "dpc_b200.h")).read()
    declared = set(re.findall(r"DPC_API[^;(]*?\b(dpc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    lib = ctypes.CDLL(dpc.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    # the ctypes prototypes cover exactly the declared surface
    assert declared == set(dpc._lib.SIGNATURES)
    assert dpc.version() == 140


def test_workspace_bytes_and_params_struct():
    import pytorch_unsup_pc_b200 as dpc
    from pytorch_unsup_pc_b200 import ops
    cfg = default_cfg(vox_size=64, vox_size_z=32)
    p = ops.make_params(cfg, 64, 8000)
    assert (p.P, p.N, p.Vz, p.V) == (64, 8000, 32, 64)
    assert p.camera_distance == 2.0 and p.focal_length == 1.875 and p.max_depth == 10.0
    assert p.drc_logsum == 1 and p.flip_y == 1
    need = dpc._lib.load().dpc_workspace_bytes(ctypes.byref(p))
    # pose partials (P*ceil(N/256)*8 doubles) + scale partials + 2*P*N sort keys
    assert need >= 64 * 32 * 8 * 8 + 2 * 64 * 8000 * 4
    assert need % 256 == 0


def test_launch_plan_queries():
    """Host-side launch plan (no GPU needed): half-batches from 64 projections on, and 6 kernel
    launches per chunk while the cloud fits the fused pose + binning cluster kernel (N <= 16384),
    7 beyond."""
    from pytorch_unsup_pc_b200 import _lib, ops
    lib = _lib.load()
    cfg = default_cfg(vox_size=64)
    plan = lambda P, N: (lib.dpc_project_chunks(ctypes.byref(ops.make_params(cfg, P, N))),
                         lib.dpc_project_kernels_per_chunk(ctypes.byref(ops.make_params(cfg, P, N))))
    assert plan(64, 8000) == (2, 6)
    assert plan(63, 8000) == (1, 6)
    assert plan(65, 1) == (1, 6)
    assert plan(128, 16384) == (2, 6)
    assert plan(4, 16385) == (1, 7)
    assert plan(4, 65535) == (1, 7)


def test_cpu_tensors_are_refused():
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.pointcloud_project_fast(cfg, torch.zeros(1, 4, 3), torch.ones(1, 4), None, None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.pc_perspective_transform(cfg, torch.zeros(1, 4, 3), torch.ones(1, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.drc_projection(torch.zeros(1, 32, 32, 32, 1), cfg)


def test_unsupported_reference_branches_raise():
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=32)
    # (all_rgb is supported: tests/test_rgb.py)
    with pytest.raises(NotImplementedError):
        dpc.pointcloud_project_fast(default_cfg(pose_quaternion=False), torch.zeros(1, 4, 3),
                                    torch.ones(1, 4, 4), None, None)
    with pytest.raises(NotImplementedError):
        dpc.pointcloud_project_fast(default_cfg(ptn_max_projection=True), torch.zeros(1, 4, 3),
                                    torch.ones(1, 4), None, None)


def test_host_taps_validation():
    from pytorch_unsup_pc_b200 import ops
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    taps = ops.host_taps(dpc.smoothing_kernel(cfg, 3.0))
    assert [t.numel() for t in taps] == [21, 21, 21]
    assert all(t.dtype == torch.float32 and not t.is_cuda for t in taps)
    assert abs(taps[0].sum().item() - 1) < 1e-6
    with pytest.raises(ValueError):
        ops.host_taps([torch.ones(4)] * 3)
    with pytest.raises(ValueError):
        ops.host_taps([torch.ones(23)] * 3)
    assert ops.host_taps(None) is None


def test_anisotropic_kernel_lengths():
    import pytorch_unsup_pc_b200 as dpc
    cfg = default_cfg(vox_size=64, vox_size_z=32, pc_gauss_kernel_size=21)
    k = dpc.smoothing_kernel(cfg, 3.0)
    assert [tuple(t.shape) for t in k] == [(1, 1, 1, 1, 21), (1, 1, 1, 21, 1), (1, 1, 11, 1, 1)]


def test_dropout_selection_is_validated():
    """A user-supplied dropout selection is range- and duplicate-checked before any kernel sees
    it (an out-of-range index would be an out-of-bounds access, a repeated one would lose a
    gradient in the backward's inverse map)."""
    from pytorch_unsup_pc_b200 import point_cloud as pc
    ok = torch.tensor([[0, 3, 2], [1, 4, 0]])
    assert pc._selection(ok, 2, 5, "cpu").dtype == torch.int32
    pairs = torch.stack([torch.zeros_like(ok), ok], dim=-1)            # the reference's [P,M,2]
    assert torch.equal(pc._selection(pairs, 2, 5, "cpu"), ok.int())
    for bad in ([[0, 3, 5], [1, 4, 0]], [[0, -1, 2], [1, 4, 0]]):
        with pytest.raises(ValueError, match="must lie in"):
            pc._selection(torch.tensor(bad), 2, 5, "cpu")
    with pytest.raises(ValueError, match="distinct"):
        pc._selection(torch.tensor([[0, 3, 3], [1, 4, 0]]), 2, 5, "cpu")
    with pytest.raises(TypeError):
        pc._selection(torch.tensor([[0.0, 1.0, 2.0], [1.0, 4.0, 0.0]]), 2, 5, "cpu")
    with pc.options(validate_indices=False):
        pc._selection(torch.tensor([[0, 3, 3], [1, 4, 0]]), 2, 5, "cpu")


def test_tap_radius_follows_sigma():
    """The radius the blur kernels run with: all but 1e-7 of the taps' mass (no GPU needed)."""
    import pytorch_unsup_pc_b200 as dpc
    lib = dpc._lib.load()
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    got = {}
    for sigma in (3.0, 2.0, 1.3, 1.0, 0.8, 0.6, 0.4, 0.2):
        taps = dpc.smoothing_kernel(cfg, sigma)[0].reshape(-1).contiguous()
        r = lib.dpc_tap_radius(taps.data_ptr(), taps.numel())
        dropped = float(taps.double()[: 10 - r].sum() + taps.double()[11 + r:].sum())
        assert dropped <= 1e-7, (sigma, r, dropped)
        if r > 0:      # one tap less would drop too much
            assert dropped + 2 * float(taps[10 - r]) > 1e-7, (sigma, r)
        got[sigma] = r
    assert got[3.0] == 10 and got[2.0] == 10 and got[1.0] <= 6 and got[0.4] <= 2 and got[0.2] == 1, got
    assert all(got[a] >= got[b] for a, b in zip(list(got)[:-1], list(got)[1:])), got
    one = torch.ones(1)
    assert lib.dpc_tap_radius(one.data_ptr(), 1) == 0
    assert lib.dpc_tap_radius(one.data_ptr(), 2) == -1
    # the run-time override (what the sigma-schedule GPU test uses instead of a child process):
    # 0 keeps every non-zero tap, a larger bound drops more, a negative value restores the default
    taps = dpc.smoothing_kernel(cfg, 1.0)[0].reshape(-1).contiguous()
    try:
        assert lib.dpc_set_tap_truncation(0.0) == 0
        nz = torch.nonzero(taps).reshape(-1)
        assert lib.dpc_tap_radius(taps.data_ptr(), taps.numel()) == int(nz.max()) - 10 == 10
        assert lib.dpc_set_tap_truncation(1e-3) == 0
        assert lib.dpc_tap_radius(taps.data_ptr(), taps.numel()) < got[1.0]
    finally:
        lib.dpc_set_tap_truncation(-1.0)
    assert lib.dpc_tap_radius(taps.data_ptr(), taps.numel()) == got[1.0]
    # (the launch-mode switch is a plain setter; its effect is covered on the GPU)
    assert lib.dpc_set_programmatic_launch(0) == 0 and lib.dpc_set_programmatic_launch(-1) == 0


def test_render_loss_argument_validation_needs_no_gpu():
    """The fused step's entry points validate geometry and pointers before any CUDA call: bad
    arguments come back as DPC_ERR_ARG with a message (no compute on the CPU box)."""
    import pytorch_unsup_pc_b200 as dpc
    from pytorch_unsup_pc_b200 import ops
    lib = dpc._lib.load()
    cfg = default_cfg(vox_size=64)
    p = ops.make_params(cfg, 64, 8000)
    taps = (None, 0, None, 0, None, 0)

    def fwd(replicas, C, G, outputs=0, gt=1):
        p.outputs = outputs
        return lib.dpc_render_loss_fwd(ctypes.byref(p), replicas, 8000, None, 1, 1, None, None, None, *taps,
                                       0, C, G, gt, None, ctypes.c_float(1.0), 1, 1, None, 1, 1, 1, 1, 1,
                                       1, 1, None, 0, None)
    assert fwd(4, 3, 128) != 0 and b"replicas" in lib.dpc_last_error()          # 4 is not views x 3
    assert fwd(4, 4, 100) != 0 and b"GT size" in lib.dpc_last_error()            # 100 is no multiple of 64
    assert fwd(4, 4, 128, outputs=1) != 0 and b"outputs" in lib.dpc_last_error()
    assert fwd(5, 5, 128) != 0                                                     # 5 does not divide 64
    assert fwd(4, 4, 128, gt=None) != 0 and b"NULL" in lib.dpc_last_error()
    # slots: winner-only where the fast ray state exists (cells, either scatter mode, cubic grid)
    p.outputs = 0
    assert lib.dpc_render_loss_slots(ctypes.byref(p), 4, 1, 0) == 16
    assert lib.dpc_render_loss_slots(ctypes.byref(p), 4, 0, 0) == 64
    assert lib.dpc_render_loss_slots(ctypes.byref(p), 4, 1, 1) == 16
    assert lib.dpc_render_loss_slots(ctypes.byref(p), 4, 0, 1) == 64
    q = ops.make_params(default_cfg(vox_size=64, vox_size_z=32), 64, 8000)
    assert lib.dpc_render_loss_slots(ctypes.byref(q), 4, 1, 0) == 64
