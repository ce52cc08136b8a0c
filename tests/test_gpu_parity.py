"""GPU parity: the CUDA path (through the Python mirror -> ctypes -> C ABI)
against the reference's golden vectors and against the CPU oracle.

Tolerances (BASELINE.json north_star; SURVEY.md section 7 hard part 2 reads
"relative" as relative to the tensor's max): forward 1e-5, gradients 1e-4.
Gradient cases use boundary-screened inputs (tests/_inputs.py).
"""
import pytest
import torch

import _golden
import _inputs
from oracle import closed_form as CF
from oracle.config import default_cfg

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 1e-4
CASE_NAMES = sorted(_golden.CASES)
GRAD_KEYS = ("points", "quat", "translation", "focal", "scale")


@pytest.fixture(scope="module")
def dpc():
    import pytorch_unsup_pc_b200 as m
    m._lib.load()          # fail loudly if the CUDA library is missing
    return m


def run_cuda(dpc, cfg, inp, P, V):
    dev = torch.device("cuda:0")
    leaves = {k: (None if inp[k] is None else inp[k].detach().to(dev).requires_grad_())
              for k in GRAD_KEYS}
    out = dpc.pointcloud_project_fast(cfg, leaves["points"], leaves["quat"], leaves["translation"],
                                      None, inp["kernel"], scaling_factor=leaves["scale"],
                                      focal_length=leaves["focal"])
    Wp, Wd = _inputs.loss_weights(P, V)
    loss = (out["proj"] * Wp.to(dev)).sum() + 0.1 * (out["proj_depth"] * Wd.to(dev)).sum()
    keys = [k for k in GRAD_KEYS if leaves[k] is not None]
    grads = dict(zip(keys, torch.autograd.grad(loss, [leaves[k] for k in keys])))
    return out, loss, grads


def run_oracle(cfg, inp, P, V):
    leaves = {k: (None if inp[k] is None else inp[k].detach().clone().requires_grad_())
              for k in GRAD_KEYS}
    out = CF.project(cfg, leaves["points"], leaves["quat"], leaves["translation"], inp["kernel"],
                     leaves["scale"], leaves["focal"])
    Wp, Wd = _inputs.loss_weights(P, V)
    loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
    keys = [k for k in GRAD_KEYS if leaves[k] is not None]
    grads = dict(zip(keys, torch.autograd.grad(loss, [leaves[k] for k in keys])))
    return out, loss, grads


@pytest.mark.parametrize("name", CASE_NAMES)
def test_golden_forward_and_gradients(dpc, name):
    """CUDA vs the vectors the REAL reference produced (tests/golden/make_golden.py)."""
    rec = _golden.load(name)
    cfg = _golden.case_cfg(name)
    inp = _golden.case_inputs(rec)
    P, V = inp["points"].shape[0], cfg.vox_size
    out, loss, grads = run_cuda(dpc, cfg, inp, P, V)
    errs = {}
    for k in ("proj", "proj_depth", "tr_pc"):
        errs[k] = _golden.rel_err(out[k], rec[k])
    for k in ("voxels", "drc_probs"):
        sub = out[k].detach().reshape(-1)[::_golden.VOX_STRIDE]
        errs[k] = _golden.rel_err(sub, rec[k + "_sub"])
    errs["loss"] = abs(loss.item() - rec["loss"]) / abs(rec["loss"])
    gerrs = {k: _golden.rel_err(g, rec["grad_" + k].reshape(g.shape)) for k, g in grads.items()}
    print(name, "fwd", {k: "%.2e" % v for k, v in errs.items()},
          "grad", {k: "%.2e" % v for k, v in gerrs.items()})
    for k, v in errs.items():
        assert v < FWD_TOL, (k, v)
    for k, v in gerrs.items():
        assert v < GRAD_TOL, (k, v)
    # the same case without the optional outputs: the ray kernels then keep their fast saved
    # state (signed clipped occupancy + transmittance checkpoints, one-sweep backward) wherever
    # the grid is cubic and the DRC is in log-sum form
    dpc.set_outputs(voxels=False, drc_probs=False)
    try:
        out2, loss2, grads2 = run_cuda(dpc, cfg, inp, P, V)
    finally:
        dpc.set_outputs(voxels=True, drc_probs=True)
    assert out2["voxels"] is None and out2["drc_probs"] is None
    for k in ("proj", "proj_depth", "tr_pc"):
        assert _golden.rel_err(out2[k], rec[k]) < FWD_TOL, k
    for k, g in grads2.items():
        assert _golden.rel_err(g, rec["grad_" + k].reshape(g.shape)) < GRAD_TOL, k


@pytest.mark.parametrize("V,K,sigma,N,P,scaled", [
    (32, 11, 1.5, 1, 1, True), (32, 3, 0.4, 37, 3, True), (32, None, None, 500, 2, False),
    (64, 5, 0.6, 1000, 3, True), (64, 21, 3.0, 4000, 2, False), (64, 11, 2.0, 333, 1, True),
    (128, 21, 3.0, 3000, 1, True), (128, 7, 1.0, 100, 2, False)])
def test_fast_ray_state_matches_general_layout(dpc, V, K, sigma, N, P, scaled):
    """Without the optional outputs the ray kernels keep the signed clipped occupancy and
    transmittance checkpoints and run the one-sweep bulk-copy backward; with them they keep B and
    run the two-sweep one.  Same mathematics: every output and gradient agrees to rounding, over
    tap radii (0, 1, 2, 5, 10 -> every ring length), grid sizes, tiny and ragged clouds."""
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=K or 11)
    case = _inputs.make_case(cfg, P, N, 31 * V + N, translation=True, focal=True, scale=scaled,
                             screened=False)
    case["kernel"] = None if K is None else CF.smoothing_taps(cfg, sigma)
    out_g, loss_g, grads_g = run_cuda(dpc, cfg, case, P, V)
    dpc.set_outputs(voxels=False, drc_probs=False)
    try:
        out_f, loss_f, grads_f = run_cuda(dpc, cfg, case, P, V)
    finally:
        dpc.set_outputs(voxels=True, drc_probs=True)
    for k in ("proj", "proj_depth"):
        assert _golden.rel_err(out_f[k], out_g[k]) < 1e-6, k
    assert torch.equal(out_f["tr_pc"], out_g["tr_pc"])
    for k in grads_g:
        assert _golden.rel_err(grads_f[k], grads_g[k]) < 2e-5, k


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_default_path_is_reproducible(dpc, kind):
    """The plane scatter accumulates in biased fixed point with integer atomics (order-free) and
    every reduction of the backward runs in a fixed order: the DEFAULT path returns the same bits
    run after run, pile-ups included."""
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    case = _inputs.make_case(cfg, 4, 8000, 515, kind=kind, translation=True, screened=False)
    case["points"][3, :3000] = case["points"][3, 0]           # 3000 points in ONE cell
    case["kernel"] = CF.smoothing_taps(cfg, 3.0)
    runs = [run_cuda(dpc, cfg, case, 4, 64) for _ in range(3)]
    for out, loss, grads in runs[1:]:
        for k in ("proj", "proj_depth", "voxels", "drc_probs", "tr_pc"):
            assert torch.equal(out[k], runs[0][0][k]), k
        for k in grads:
            assert torch.equal(grads[k], runs[0][2][k]), k


@pytest.mark.parametrize("outputs", [True, False])
def test_half_batches_match_separate_calls(dpc, outputs):
    """P >= 64 runs as two half-batches on two internal streams (dpc_project_chunks): every
    projection must come out exactly as if its half had been projected alone -- bit for bit in
    the deterministic mode, to scatter-order rounding otherwise -- in both saved-state layouts."""
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    P, N = 64, 700
    case = _inputs.make_case(cfg, P, N, 4242, translation=True, scale=True, screened=False)
    kern = CF.smoothing_taps(cfg, 1.5)
    Wp, Wd = (w.to(dev) for w in _inputs.loss_weights(P, 32))

    def run(sl):
        leaves = [case[k][sl].to(dev).requires_grad_() for k in ("points", "quat", "translation", "scale")]
        out = dpc.pointcloud_project_fast(cfg, leaves[0], leaves[1], leaves[2], None, kern,
                                          scaling_factor=leaves[3])
        loss = (out["proj"] * Wp[sl]).sum() + 0.1 * (out["proj_depth"] * Wd[sl]).sum()
        return out, torch.autograd.grad(loss, leaves)

    dpc.set_outputs(voxels=outputs, drc_probs=outputs)
    try:
        for det in (True, False):
            dpc.set_deterministic(det)
            whole, gw = run(slice(0, P))
            for sl in (slice(0, P // 2), slice(P // 2, P)):
                part, gp = run(sl)
                for k in ("proj", "proj_depth", "tr_pc"):
                    if det or k == "tr_pc":
                        assert torch.equal(whole[k][sl], part[k]), (det, k)
                    else:
                        assert _golden.rel_err(whole[k][sl], part[k]) < FWD_TOL, (det, k)
                for a, b in zip(gw, gp):
                    if det:
                        assert torch.equal(a[sl], b)
                    else:
                        assert _golden.rel_err(a[sl], b) < GRAD_TOL
    finally:
        dpc.set_deterministic(False)
        dpc.set_outputs(voxels=True, drc_probs=True)


@pytest.mark.parametrize("seed,sigma,kind", [(11, 3.0, "uniform"), (12, 0.7, "uniform"),
                                             (13, 1.5, "clustered")])
def test_oracle_fresh_seeds(dpc, seed, sigma, kind):
    """CUDA vs the oracle on inputs no fixture has seen."""
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    case = _inputs.make_case(cfg, 4, 1500, seed, kind=kind, translation=(seed % 2 == 1),
                             focal=(seed % 2 == 0), screened=True)
    case["kernel"] = CF.smoothing_taps(cfg, sigma)
    o_out, o_loss, o_grads = run_oracle(cfg, case, 4, 32)
    c_out, c_loss, c_grads = run_cuda(dpc, cfg, case, 4, 32)
    for k in ("proj", "proj_depth", "tr_pc", "voxels", "drc_probs"):
        assert _golden.rel_err(c_out[k], o_out[k]) < FWD_TOL, k
    for k in o_grads:
        assert _golden.rel_err(c_grads[k], o_grads[k]) < GRAD_TOL, k


def test_deterministic_mode_is_bit_exact_and_close(dpc):
    """Config 5: sort-then-segment scatter, run twice, torch.equal on everything."""
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    case = _inputs.make_case(cfg, 4, 8000, 1005, kind="clustered", screened=False)
    case["kernel"] = CF.smoothing_taps(cfg, 3.0)
    runs = []
    with dpc.options(deterministic=True):
        for _ in range(2):
            out, loss, grads = run_cuda(dpc, cfg, case, 4, 64)
            runs.append((out, loss, grads))
    for k in ("proj", "proj_depth", "tr_pc", "voxels", "drc_probs"):
        assert torch.equal(runs[0][0][k], runs[1][0][k]), k
    for k in runs[0][2]:
        assert torch.equal(runs[0][2][k], runs[1][2][k]), k
    assert runs[0][1].item() == runs[1][1].item()
    # and it agrees with the default (atomic) scatter to rounding
    out_a, _, grads_a = run_cuda(dpc, cfg, case, 4, 64)
    for k in ("proj", "proj_depth", "voxels"):
        assert _golden.rel_err(out_a[k], runs[0][0][k]) < FWD_TOL, k
    for k in grads_a:
        assert _golden.rel_err(grads_a[k], runs[0][2][k]) < GRAD_TOL, k


def test_frustum_edges(dpc):
    """Coordinates exactly on the frustum faces and outside it.

    Identity pose: tr_pc = (p0, f*p1/(p0+d), f*p2/(p0+d)).  A coordinate of
    exactly -0.5 is valid (cell 0); exactly +0.5 makes the reference index
    cell V and raise -- the CUDA path drops the
    zero-weight out-of-range corners instead (oracle: drop_oob=True)."""
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    pts = torch.tensor([[[0.5, 0.0, 0.0], [-0.5, 0.0, 0.0], [0.6, 0.0, 0.0], [0.0, 0.9, 0.0],
                         [0.25, 0.1, -0.2], [-0.5, -0.4, -0.4], [0.0, 0.0, 0.0]]])
    quat = torch.tensor([[2.0, 0.0, 0.0, 0.0]])
    orc = CF.project(cfg, pts, quat, None, CF.smoothing_taps(cfg, 1.0), None, drop_oob=True)
    dev = torch.device("cuda:0")
    out = dpc.pointcloud_project_fast(cfg, pts.to(dev), quat.to(dev), None, None,
                                      CF.smoothing_taps(cfg, 1.0))
    for k in ("proj", "proj_depth", "tr_pc", "voxels", "drc_probs"):
        assert _golden.rel_err(out[k], orc[k]) < FWD_TOL, k
    # stand-alone scatter of the fp32 tr_pc (rounding to fp32 can move a point
    # onto the face, so the oracle scatters the same fp32 values)
    tr32 = orc["tr_pc"].float()
    raw, _ = dpc.pointcloud2voxels3d_fast(cfg, tr32.to(dev), None)
    assert _golden.rel_err(raw, CF.scatter_trilinear(cfg, tr32.double(), drop_oob=True)) < FWD_TOL
    # mass conservation: every in-frustum point deposits total weight 1
    inside = ((tr32 >= -0.5) & (tr32 <= 0.5)).all(-1).sum().item()
    assert abs(raw.sum().item() - inside) < 1e-4


def test_all_points_outside_and_single_point(dpc):
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    dev = torch.device("cuda:0")
    quat = torch.tensor([[1.0, 0.0, 0.0, 0.0]], device=dev)
    kern = CF.smoothing_taps(cfg, 1.0)
    far = torch.full((1, 5, 3), 3.0, device=dev, requires_grad=True)
    out = dpc.pointcloud_project_fast(cfg, far, quat, None, None, kern)
    orc = CF.project(cfg, far.detach().cpu(), quat.cpu(), None, kern, None)
    assert _golden.rel_err(out["proj"], orc["proj"]) < FWD_TOL
    assert _golden.rel_err(out["proj_depth"], orc["proj_depth"]) < FWD_TOL
    (g,) = torch.autograd.grad(out["proj"].sum(), [far])
    assert torch.count_nonzero(g).item() == 0
    one = torch.tensor([[[0.1, -0.2, 0.05]]], device=dev)
    out1 = dpc.pointcloud_project_fast(cfg, one, quat, None, None, kern)
    orc1 = CF.project(cfg, one.cpu(), quat.cpu(), None, kern, None)
    for k in ("proj", "proj_depth", "voxels"):
        assert _golden.rel_err(out1[k], orc1[k]) < FWD_TOL, k


def test_upstream_grads_on_every_output(dpc):
    """Gradients flowing in through voxels, drc_probs and tr_pc as well."""
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    case = _inputs.make_case(cfg, 2, 700, 21, translation=True, screened=True)
    kern = CF.smoothing_taps(cfg, 1.5)
    g = torch.Generator().manual_seed(3)
    Wv = torch.rand(2, 32, 32, 32, 1, generator=g)
    Wq = torch.rand(33, 2, 32, 32, 1, generator=g)
    Wt = torch.rand(2, 700, 3, generator=g)

    def loss_of(out, dev):
        return ((out["voxels"] * Wv.to(dev)).sum() + (out["drc_probs"] * Wq.to(dev)).sum()
                + (out["tr_pc"] * Wt.to(dev)).sum() + out["proj"].sum())

    leaves_o = {k: case[k].clone().requires_grad_() for k in ("points", "quat", "translation", "scale")}
    o = CF.project(cfg, leaves_o["points"], leaves_o["quat"], leaves_o["translation"], kern,
                   leaves_o["scale"])
    go = torch.autograd.grad(loss_of(o, "cpu"), list(leaves_o.values()))
    dev = torch.device("cuda:0")
    leaves_c = {k: case[k].to(dev).requires_grad_() for k in leaves_o}
    c = dpc.pointcloud_project_fast(cfg, leaves_c["points"], leaves_c["quat"],
                                    leaves_c["translation"], None, kern,
                                    scaling_factor=leaves_c["scale"])
    gc = torch.autograd.grad(loss_of(c, dev), list(leaves_c.values()))
    for k, a, b in zip(leaves_o, gc, go):
        assert _golden.rel_err(a, b) < GRAD_TOL, k


@pytest.mark.parametrize("V,N,sigma,kind", [(64, 8000, 3.0, "uniform"), (32, 1500, 0.7, "clustered"),
                                            (128, 4000, 3.0, "uniform")])
def test_plane_local_path_matches_global_grid_path(dpc, V, N, sigma, kind):
    """The default path (planes built / gathered in shared memory) against the
    path that keeps the raw grid in global memory: same math, different
    association of the fp32 sums -> agreement to rounding, on every output
    and gradient, including a degenerate cloud that piles into few cells."""
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=21 if V > 32 else 11)
    case = _inputs.make_case(cfg, 3, N, 77 + V, kind=kind, translation=True, focal=True,
                             screened=False)
    case["points"][2, : N // 2] = case["points"][2, 0]          # N/2 points in ONE cell
    case["kernel"] = CF.smoothing_taps(cfg, sigma)
    out_p, loss_p, grads_p = run_cuda(dpc, cfg, case, 3, V)
    with dpc.options(plane_local=False):
        out_g, loss_g, grads_g = run_cuda(dpc, cfg, case, 3, V)
    for k in ("proj", "proj_depth", "voxels", "drc_probs"):
        assert _golden.rel_err(out_p[k], out_g[k]) < FWD_TOL, k
    assert torch.equal(out_p["tr_pc"], out_g["tr_pc"])
    for k in grads_p:
        assert _golden.rel_err(grads_p[k], grads_g[k]) < GRAD_TOL, k


@pytest.mark.parametrize("V,N,sigma,kind", [(64, 8000, 3.0, "uniform"), (32, 1500, 0.7, "clustered"),
                                            (128, 4000, 3.0, "uniform"), (64, 300, 1.0, "clustered")])
def test_sorted_plane_build_equals_standalone_sorted_scatter(dpc, V, N, sigma, kind):
    """Deterministic mode: the plane kernel sums every plane row from the row-sorted records in
    the order of the stand-alone sort-then-segment scatter (which writes the raw grid to global
    memory), so every forward output is BIT-identical between the two; the backward (plane
    gather vs grid gather) agrees to rounding.  Includes a pile of points in one cell (raw > 1:
    the clamp and its gate bits) and a translated, focal-scaled pose (points leave the frustum)."""
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=21 if V > 32 else 11)
    case = _inputs.make_case(cfg, 3, N, 177 + V, kind=kind, translation=True, focal=True,
                             screened=False)
    case["points"][2, : N // 2] = case["points"][2, 0]          # N/2 points in ONE cell
    case["kernel"] = CF.smoothing_taps(cfg, sigma)
    with dpc.options(deterministic=True):
        out_p, loss_p, grads_p = run_cuda(dpc, cfg, case, 3, V)
        out_p2, _, grads_p2 = run_cuda(dpc, cfg, case, 3, V)
        with dpc.options(plane_local=False):
            out_g, loss_g, grads_g = run_cuda(dpc, cfg, case, 3, V)
    for k in ("proj", "proj_depth", "tr_pc", "voxels", "drc_probs"):
        assert torch.equal(out_p[k], out_g[k]), k
        assert torch.equal(out_p[k], out_p2[k]), k
    for k in grads_p:
        assert torch.equal(grads_p[k], grads_p2[k]), k
        assert _golden.rel_err(grads_p[k], grads_g[k]) < GRAD_TOL, k


def test_release_and_recreate_internal_streams(dpc):
    """dpc_release() destroys the calling thread's internal side streams / events (the only state
    the library keeps); the next batch of >= 64 projections re-creates them and gives the same bits."""
    import ctypes
    from pytorch_unsup_pc_b200 import _lib, ops
    lib = _lib.load()
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    P, N = 64, 500
    case = _inputs.make_case(cfg, P, N, 77, scale=True, screened=False)
    taps = ops.host_taps(CF.smoothing_taps(cfg, 1.5))
    params = ops.make_params(cfg, P, N, flip_y=True)
    d = {k: case[k].to(dev).contiguous() for k in ("points", "quat", "scale")}
    f32 = dict(dtype=torch.float32, device=dev)
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    cells = torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def forward():
        grid, bits = torch.empty(P, 32, 32, 32, **f32), torch.empty(P, 32, 32, 1, dtype=torch.int32, device=dev)
        mask, depth, tr = torch.empty(P, 32, 32, **f32), torch.empty(P, 32, 32, **f32), torch.empty(P, N, 3, **f32)
        _lib.check(lib.dpc_project_fwd(ctypes.byref(params), d["points"].data_ptr(), d["quat"].data_ptr(),
                                       None, None, d["scale"].data_ptr(), *ops._tap_args(taps), 0,
                                       tr.data_ptr(), grid.data_ptr(), bits.data_ptr(), cells.data_ptr(),
                                       mask.data_ptr(), depth.data_ptr(), None, None, ws.data_ptr(),
                                       ws.numel(), sp), "fwd")
        torch.cuda.synchronize(dev)
        return mask, depth, tr
    assert lib.dpc_project_chunks(ctypes.byref(params)) == 2
    a = forward()
    assert lib.dpc_release() == 0
    assert lib.dpc_release() == 0            # nothing left to release: still fine
    b = forward()
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_capped_chunks_match_separate_calls(dpc):
    """At 64^3 a batch above 128 projections runs as chunks of 64 (64 MiB of grid, api.cu
    chunk_size) -- 130 projections: 64 + 64 + 2, alternating over the two internal streams.  Every
    projection must come out as if its chunk had been projected alone."""
    import ctypes
    from pytorch_unsup_pc_b200 import _lib, ops
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=5)
    P, N = 130, 200
    assert _lib.load().dpc_project_chunks(ctypes.byref(ops.make_params(cfg, P, N))) == 3
    case = _inputs.make_case(cfg, P, N, 777, scale=True, screened=False)
    kern = CF.smoothing_taps(cfg, 0.8)
    Wp, Wd = (w.to(dev) for w in _inputs.loss_weights(P, 64))

    def run(sl):
        leaves = [case[k][sl].to(dev).requires_grad_() for k in ("points", "quat", "scale")]
        out = dpc.pointcloud_project_fast(cfg, leaves[0], leaves[1], None, None, kern,
                                          scaling_factor=leaves[2])
        loss = (out["proj"] * Wp[sl]).sum() + 0.1 * (out["proj_depth"] * Wd[sl]).sum()
        return out, torch.autograd.grad(loss, leaves)

    dpc.set_outputs(voxels=False, drc_probs=False)
    try:
        whole, gw = run(slice(0, P))
        for sl in (slice(0, 64), slice(64, 128), slice(128, 130)):
            part, gp = run(sl)
            for k in ("proj", "proj_depth", "tr_pc"):
                assert torch.equal(whole[k][sl], part[k]), k      # order-free scatter: bit-equal
            for a, b in zip(gw, gp):
                assert torch.equal(a[sl], b)
    finally:
        dpc.set_outputs(voxels=True, drc_probs=True)
