#!/usr/bin/env python
"""The timed loop of bench.py's `value` leg and nothing else: three fwd+bwd steps (rotating input
sets) captured into one CUDA graph and replayed -- the command ncu wraps for STEADY-STATE cache
behaviour (`--cache-control none --graph-profiling node`): half-batch launches, L2 contents as the
previous kernels of the chain left them.     python scripts/profile_graph.py [--replays 4]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import pytorch_unsup_pc_b200 as dpc  # noqa: E402
from pytorch_unsup_pc_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="A")
    ap.add_argument("--replays", type=int, default=4)
    a = ap.parse_args()
    lib = _lib.load()
    w = bench.WORKLOADS[a.workload]
    cfg = bench.make_cfg(w)
    dev = torch.device("cuda:0")
    P, N, V = w["P"], w["N"], w["V"]
    params = ops.make_params(cfg, P, N, flip_y=True)
    taps_h = ops.host_taps(dpc.smoothing_kernel(cfg, w["sigma"]))
    taps = ops._tap_args(taps_h)
    sets = [{k: v.to(dev) for k, v in bench.synth_inputs(w, 1000 + i).items()} for i in range(3)]
    f32 = dict(dtype=torch.float32, device=dev)
    tr_pc, grid, g_grid = torch.empty(P, N, 3, **f32), torch.empty(P, V, V, V, **f32), torch.empty(P, V, V, V, **f32)
    bits = torch.empty(P, V, V, V // 32, dtype=torch.int32, device=dev)
    mask, depth = torch.empty(P, V, V, **f32), torch.empty(P, V, V, **f32)
    g_points, g_quat, g_scale = torch.empty(P, N, 3, **f32), torch.empty(P, 4, **f32), torch.empty(P, **f32)
    cells = torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    p_ = ops._ptr

    def step(d):
        _lib.check(lib.dpc_project_fwd(ctypes.byref(params), p_(d["points"]), p_(d["quat"]), None, None,
                                       p_(d["scale"]), *taps, _lib.SCATTER_ATOMIC, p_(tr_pc), p_(grid),
                                       p_(bits), p_(cells), p_(mask), p_(depth), None, None, p_(ws),
                                       ws.numel(), sp), "fwd")
        _lib.check(lib.dpc_project_bwd(ctypes.byref(params), p_(d["points"]), p_(d["quat"]), None, None,
                                       p_(d["scale"]), *taps, p_(grid), p_(bits), p_(cells), p_(d["g_mask"]),
                                       p_(d["g_depth"]), None, None, None, p_(g_grid), p_(g_points), p_(g_quat),
                                       None, None, p_(g_scale), p_(ws), ws.numel(), sp), "bwd")
    for d in sets:
        step(d)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    cap.wait_stream(stream)
    with torch.cuda.stream(cap):
        sp.value = cap.cuda_stream
        with torch.cuda.graph(g, stream=cap):
            for d in sets:
                step(d)
        sp.value = stream.cuda_stream
    stream.wait_stream(cap)
    for _ in range(a.replays):
        g.replay()
    torch.cuda.synchronize()
    print("replayed", a.replays, "x 3 steps; checksum", float(mask.sum()), float(g_points.abs().sum()))


if __name__ == "__main__":
    main()
