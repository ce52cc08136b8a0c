"""Dev helper: step time along the training trajectory -- sigma_rel 3.0 -> 0.2 (tap radius 10 -> 1)
x point-dropout keep probability 0.07 -> 1.0 (N_eff 560 -> 8000) -- of the plain fwd+bwd step
(C ABI, graph replay) at workload A shapes.  python scripts/sigma_sweep.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import _lib, ops


def step_us(sigma, keep, iters=60):
    w = dict(bench.WORKLOADS["A"], sigma=sigma, N=max(1, int(8000 * keep)))
    cfg = bench.make_cfg(w)
    lib = _lib.load()
    dev = torch.device("cuda:0")
    P, N, V = w["P"], w["N"], w["V"]
    taps = ops.host_taps(dpc.smoothing_kernel(cfg, sigma))
    params = ops.make_params(cfg, P, N, flip_y=True)
    d = {k: v.to(dev) for k, v in bench.synth_inputs(w, 1000).items()}
    f32 = dict(dtype=torch.float32, device=dev)
    buf = dict(tr_pc=torch.empty(P, N, 3, **f32), grid=torch.empty(P, V, V, V, **f32),
               bits=torch.empty(P, V, V, V // 32, dtype=torch.int32, device=dev),
               mask=torch.empty(P, V, V, **f32), depth=torch.empty(P, V, V, **f32),
               g_grid=torch.empty(P, V, V, V, **f32), g_points=torch.empty(P, N, 3, **f32),
               g_quat=torch.empty(P, 4, **f32), g_scale=torch.empty(P, **f32),
               cells=torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev))
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    P_ = ops._ptr
    ta = ops._tap_args(taps)

    def step():
        _lib.check(lib.dpc_project_fwd(ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None,
                                       P_(d["scale"]), *ta, 0, P_(buf["tr_pc"]), P_(buf["grid"]), P_(buf["bits"]),
                                       P_(buf["cells"]), P_(buf["mask"]), P_(buf["depth"]), None, None, P_(ws),
                                       ws.numel(), sptr), "fwd")
        _lib.check(lib.dpc_project_bwd(ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None,
                                       P_(d["scale"]), *ta, P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]),
                                       P_(d["g_mask"]), P_(d["g_depth"]), None, None, None, P_(buf["g_grid"]),
                                       P_(buf["g_points"]), P_(buf["g_quat"]), None, None, P_(buf["g_scale"]),
                                       P_(ws), ws.numel(), sptr), "bwd")
    step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    cap.wait_stream(stream)
    with torch.cuda.stream(cap):
        sptr.value = cap.cuda_stream
        with torch.cuda.graph(g, stream=cap):
            for _ in range(3):
                step()
        sptr.value = stream.cuda_stream
    stream.wait_stream(cap)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters // 3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = taps[0]
    return e0.elapsed_time(e1) / (iters // 3 * 3) * 1e3, lib.dpc_tap_radius(t.data_ptr(), t.numel())


if __name__ == "__main__":
    for keep in (1.0, 0.5, 0.07):
        for sigma in (3.0, 2.0, 1.3, 1.0, 0.8, 0.5, 0.2):
            us, r = step_us(sigma, keep)
            print("keep %.2f sigma %.1f radius %2d: %.1f us/step  %.0f proj/s" % (keep, sigma, r, us, 64 / us * 1e6), flush=True)
