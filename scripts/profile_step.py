#!/usr/bin/env python
"""Minimal fwd+bwd loop of the projection through the C ABI (device-resident
inputs, no CUDA graph) -- the command ncu wraps (B200_PROFILING.md):

    python scripts/profile_step.py [--workload A|B] [--steps 3] [--global-grid] [--P n] [--N n]

Each step launches, in order: pose_bin (pose + z-binning), blur_xy (plane scatter),
blurz_drc_fwd, drc_blurz_bwd_fast, blur_xy (plane gather), gather_pose_bwd -> 6 kernels per
chunk and step (DPC_CHUNK=<P or more>: one chunk).
"""
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--global-grid", action="store_true", help="memset + global scatter path")
    ap.add_argument("--P", type=int, default=0, help="override the workload's projection count")
    ap.add_argument("--N", type=int, default=0, help="override the workload's point count")
    a = ap.parse_args()
    lib = _lib.load()
    w = dict(bench.WORKLOADS[a.workload])
    if a.P:
        w["P"] = a.P
    if a.N:
        w["N"] = a.N
    cfg = bench.make_cfg(w)
    dev = torch.device("cuda:0")
    P, N, V = w["P"], w["N"], w["V"]
    params = ops.make_params(cfg, P, N, flip_y=True)
    host_taps = ops.host_taps(dpc.smoothing_kernel(cfg, w["sigma"]))   # must outlive the calls
    taps = ops._tap_args(host_taps)
    d = {k: v.to(dev) for k, v in bench.synth_inputs(w, 1000).items()}
    f32 = dict(dtype=torch.float32, device=dev)
    u8 = dict(dtype=torch.uint8, device=dev)
    tr_pc, grid, g_grid = torch.empty(P, N, 3, **f32), torch.empty(P, V, V, V, **f32), torch.empty(P, V, V, V, **f32)
    bits = torch.empty(P, V, V, V // 32, dtype=torch.int32, device=dev)
    mask, depth = torch.empty(P, V, V, **f32), torch.empty(P, V, V, **f32)
    g_points, g_quat, g_scale = torch.empty(P, N, 3, **f32), torch.empty(P, 4, **f32), torch.empty(P, **f32)
    cells = None if a.global_grid else torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), **u8)
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), **u8)
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p_ = ops._ptr
    for _ in range(a.steps):
        _lib.check(lib.dpc_project_fwd(ctypes.byref(params), p_(d["points"]), p_(d["quat"]), None, None,
                                       p_(d["scale"]), *taps, _lib.SCATTER_ATOMIC, p_(tr_pc), p_(grid),
                                       p_(bits), p_(cells), p_(mask), p_(depth), None, None, p_(ws),
                                       ws.numel(), sp), "fwd")
        _lib.check(lib.dpc_project_bwd(ctypes.byref(params), p_(d["points"]), p_(d["quat"]), None, None,
                                       p_(d["scale"]), *taps, p_(grid), p_(bits), p_(cells),
                                       p_(d["g_mask"]), p_(d["g_depth"]), None, None, None, p_(g_grid),
                                       p_(g_points), p_(g_quat), None, None, p_(g_scale), p_(ws),
                                       ws.numel(), sp), "bwd")
    torch.cuda.synchronize()
    print("ok", float(mask.sum()), float(g_points.abs().sum()))


if __name__ == "__main__":
    main()
