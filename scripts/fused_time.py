"""Dev helper: where the fused renderer + loss step's time goes -- the forward call, the backward
call and the whole step, each captured into its own CUDA graph and replayed (C ABI,
device-resident inputs).   python scripts/fused_time.py [A|C3|B] [iters]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import _lib, ops


def main(workload="A", iters=200):
    w = bench.WORKLOADS[workload]
    cfg = bench.make_cfg(w)
    lib = _lib.load()
    dev = torch.device("cuda:0")
    P, N, V, C, G = w["P"], w["N"], w["V"], w["cands"], w["G"]
    R = w["views"] * C
    B, BV = P // R, P // C
    taps = ops.host_taps(dpc.smoothing_kernel(cfg, w["sigma"]))
    params = ops.make_params(cfg, P, N, flip_y=True)
    h = bench.synth_inputs(w, 1000)
    d = {"points": h["points"][::R].contiguous().to(dev), "quat": h["quat"].to(dev),
         "scale": h["scale"].to(dev), "masks": bench.synth_masks(w, 1000).to(dev)}
    f32 = dict(dtype=torch.float32, device=dev)
    slots = lib.dpc_render_loss_slots(ctypes.byref(params), C, 1, 0)
    buf = dict(grid=torch.empty(P, V, V, V, **f32), bits=torch.empty(P, V, V, V // 32, dtype=torch.int32, device=dev),
               cells=torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev),
               mask=torch.empty(P, V, V, **f32), all_loss=torch.empty(BV, C, **f32),
               min_idx=torch.empty(BV, dtype=torch.int64, device=dev), view_loss=torch.empty(BV, **f32),
               loss=torch.empty(1, **f32), winners=torch.empty(BV, dtype=torch.int32, device=dev),
               kcoef=torch.empty(BV, **f32), g_grid=torch.empty(slots, V, V, V, **f32),
               g_rep=torch.empty(slots, N, 3, **f32), g_points=torch.empty(B, N, 3, **f32),
               g_quat=torch.empty(P, 4, **f32), g_scale=torch.empty(P, **f32))
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    P_ = ops._ptr
    ta = ops._tap_args(taps)

    def fwd():
        _lib.check(lib.dpc_render_loss_fwd(
            ctypes.byref(params), R, N, None, P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]),
            *ta, 0, C, G, P_(d["masks"]), None, ctypes.c_float(1.0), P_(buf["grid"]), P_(buf["bits"]),
            P_(buf["cells"]), P_(buf["mask"]), P_(buf["all_loss"]), P_(buf["min_idx"]), P_(buf["view_loss"]),
            P_(buf["loss"]), P_(buf["winners"]), P_(buf["kcoef"]), P_(ws), ws.numel(), sptr), "fwd")

    def bwd():
        _lib.check(lib.dpc_render_loss_bwd(
            ctypes.byref(params), R, N, None, P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]),
            *ta, 0, C, G, P_(d["masks"]), None, ctypes.c_float(1.0), P_(buf["grid"]), P_(buf["bits"]),
            P_(buf["cells"]), P_(buf["mask"]), P_(buf["min_idx"]), P_(buf["winners"]), P_(buf["kcoef"]), None,
            P_(buf["g_grid"]), P_(buf["g_rep"]), None, None, P_(buf["g_points"]), P_(buf["g_quat"]), None,
            None, P_(buf["g_scale"]), P_(ws), ws.numel(), sptr), "bwd")

    def both():
        fwd()
        bwd()

    def graph_time(fn, reps):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(stream)
        with torch.cuda.stream(cap):
            saved = sptr.value
            sptr.value = cap.cuda_stream
            with torch.cuda.graph(g, stream=cap):
                for _ in range(reps):
                    fn()
            sptr.value = saved
        stream.wait_stream(cap)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(1, iters // reps)
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * reps) * 1e3
    print(os.path.basename(_lib.LIB_PATH), workload, "slots", slots,
          "fwd %.1f us  bwd %.1f us  step %.1f us" % (graph_time(fwd, 3), graph_time(bwd, 3), graph_time(both, 3)))


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["A"]), *[int(x) for x in sys.argv[2:3]])
