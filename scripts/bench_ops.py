#!/usr/bin/env python
"""Measurement of the widened rows: f1 candidate-selection loss,
f4 Chamfer nearest neighbour.  CUDA events, inputs resident, CPU oracle beside.

    python scripts/bench_ops.py          # prints one JSON line per op
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import pytorch_unsup_pc_b200 as dpc  # noqa: E402
from oracle import chamfer as OC  # noqa: E402
from oracle import loss as OL  # noqa: E402
from oracle.config import default_cfg  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6551.7
if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def cuda_ms(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_loss():
    # config 3 per GPU: batch 16 x 4 views = 64 views x 4 candidates = 256 masks, 128^2 GT -> 64^2
    BV, C, V, G = 64, 4, 64, 128
    cfg = default_cfg(pose_predict_num_candidates=C)
    g = torch.Generator().manual_seed(1)
    masks = (torch.rand(BV, 1, G, G, generator=g) > 0.5).float()
    projs = torch.rand(BV * C, V, V, 1, generator=g)
    dm, dp = masks.to(dev), projs.to(dev).requires_grad_()

    def step():
        total, _ = dpc.add_proj_loss(cfg, {"masks": dm}, {"projs": dp}, 1.0)
        return torch.autograd.grad(total, dp)
    ms = cuda_ms(step)
    # the eager call is bound by the host (two autograd ops ~ 0.17 ms of Python for ~15 us of
    # kernels): the same step replayed from a CUDA graph gives the device time
    gs = dpc.GraphedSteps(lambda k: step(), 10, dev)
    ms_graph = cuda_ms(gs.replay, iters=20) / 10
    # algorithmic bytes: fwd reads gt + pred, bwd reads gt + pred and writes g_pred
    nbytes = 2 * (BV * G * G * 4) + 3 * (BV * C * V * V * 4)
    t0 = time.perf_counter()
    p = projs.double().requires_grad_()
    total, _ = OL.add_proj_loss(masks, p, C, 1.0)
    torch.autograd.grad(total, p)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"op": "candidate_loss fwd+bwd (python API)", "BV": BV, "C": C, "V": V, "G": G,
                      "ms_eager": ms, "ms": ms_graph, "algorithmic_bytes": nbytes,
                      "GB/s": nbytes / ms_graph / 1e6,
                      "frac_of_hbm_peak": nbytes / ms_graph / 1e6 / PEAK, "cpu_oracle_ms": cpu_ms,
                      "cpu_threads": torch.get_num_threads()}))


def bench_chamfer():
    N, M = 8000, 100000       # predicted cloud vs a dense ground-truth cloud
    g = torch.Generator().manual_seed(2)
    Vs, Vt = torch.rand(N, 3, generator=g), torch.rand(M, 3, generator=g)
    ds, dt = Vs.to(dev), Vt.to(dev)
    ms = cuda_ms(lambda: dpc.point_cloud_distance(ds, dt), iters=20)
    pairs = N * M
    t0 = time.perf_counter()
    OC.point_cloud_distance(Vs.numpy()[:500], Vt.numpy())
    cpu_s = (time.perf_counter() - t0) * N / 500
    # 8 fp32 ops + compare per pair
    print(json.dumps({"op": "point_cloud_distance", "N": N, "M": M, "ms": ms,
                      "Gpairs/s": pairs / ms / 1e6, "fp32_TFLOP/s": 8 * pairs / ms / 1e9,
                      "cpu_oracle_s_extrapolated": cpu_s}))


if __name__ == "__main__":
    bench_loss()
    bench_chamfer()
