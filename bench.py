#!/usr/bin/env python
"""bench.py -- fwd+bwd point-cloud projections/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload A|B|C3|C5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port on host cores)

A step = one forward + backward pass of the projection path over one batch of
synthetic projections (workload A: P = 16 x 4 pose candidates = 64 clouds of
8000 points -> 64^3 grid -> 64^2 mask + depth, K=21, sigma=3; BASELINE.json
configs[1]).  One process per GPU, projections sharded by (batch x candidate)
with no data-path collective (weak scaling: P per GPU is fixed).

The top-level keys are workload A's.  The default invocation also runs the other
BASELINE.json configurations, bounded, and reports them under `workloads`:
  C3  configs[2], the projection of the train step per GPU (256 projections, 64^3)
  B   configs[3], paper scale (128 projections of 16000 points, 128^3)
  C5  configs[4], depth path + deterministic sort-then-segment scatter (+ bit-exact repeat)
  train3  configs[2] itself: CNN encoder + decoder + pose ensemble + projection + loss + Adam
          under DistributedDataParallel (NCCL all-reduce of the weight gradients)
each with the same keys (`value`, `e2e`, `roofline`, `parity_check`).

`value`   device-resident inputs, the two C-ABI calls dpc_project_fwd /
          dpc_project_bwd per step, timed with CUDA events over exactly K steps
          (barrier + synchronize on both sides, max over ranks).
`e2e`     the same metric through the public Python API
          (pointcloud_project_fast + autograd) with pinned HOST inputs copied
          in and the results (mask, depth, gradients) copied out every step;
          `e2e.link_ceiling` = the rate at which this box moves the step's bytes
          (both directions at once, all ranks at once), measured in the same run.
`fused`   the renderer + candidate-selection loss as one step
          (project_candidates_loss: clouds, poses, ground-truth masks in; loss and
          cloud / pose / scale gradients out), device-resident (`value`) and end to end.
`parity_check`  after the timed region the buffers of the LAST timed step are compared with
                the CPU oracle on the same inputs for a sample of projections from both
                half-batches (forward 1e-5, gradients 1e-4, scale-relative).
`roofline`      the longest kernel of the step, from per-stage CUDA-event
                timings taken in this run (dpc_project_profile): contract bytes against the
                HBM peak (`frac`), and its FMA count against the FFMA2 rate measured on this
                GPU in this run (`fp32_frac`, dpc_fma_rate_probe).
`cpu_baseline`  the oracle port (the reference's algorithm with the
                reference's torch ops, fp64) timed on this host's cores on a
                bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fwd+bwd point-cloud projections/sec (8k pts, 64^3 grid)"
UNIT = "projections/s"

WORKLOADS = {
    # BASELINE.json configs[1]: projection-only microbench
    "A": dict(name="projection microbench: batch 16 x 4 pose candidates, 8000 pts -> 64^3 -> "
                   "64^2 mask+depth, K=21 sigma=3.0, fwd+bwd",
              P=64, N=8000, V=64, K=21, sigma=3.0, clouds=16, views=1, cands=4, G=128),
    # BASELINE.json configs[3]: paper scale
    "B": dict(name="paper scale: batch 32 x 4 candidates, 16000 pts -> 128^3 -> 128^2, K=21 "
                   "sigma=3.0, fwd+bwd",
              P=128, N=16000, V=128, K=21, sigma=3.0, clouds=32, views=1, cands=4, G=128),
    # BASELINE.json configs[2], the projection part per GPU: batch 16 x 4 views x 4 candidates
    "C3": dict(name="chair_unsupervised train-step shapes per GPU: batch 16 x 4 views x 4 candidates, "
                    "8000 pts -> 64^3 -> 64^2, K=21 sigma=3.0, fwd+bwd",
               P=256, N=8000, V=64, K=21, sigma=3.0, clouds=16, views=4, cands=4, G=128),
    # BASELINE.json configs[4]: chair_camera_supervision (one pose per view, the GT camera) through
    # mask AND depth, deterministic sort-then-segment scatter
    "C5": dict(name="chair_camera_supervision depth path: batch 16 x 4 views x 1 pose, 8000 pts -> "
                    "64^3 -> 64^2 mask+depth, K=21 sigma=3.0, deterministic sort-then-segment "
                    "scatter, fwd+bwd",
               P=64, N=8000, V=64, K=21, sigma=3.0, clouds=16, views=4, cands=1, G=128,
               deterministic=True),
}
N_INPUT_SETS = 3          # distinct input sets rotated between steps
E2E_GRAPH_STEPS = int(os.environ.get("DPC_E2E_GRAPH_STEPS", "24"))   # e2e steps captured per CUDA graph (a multiple of N_INPUT_SETS)
FWD_TOL, GRAD_TOL = 1e-5, 1e-4


def algorithmic_bytes(N, V, Vz):
    """14 G + 72 N + 4 I per fwd+bwd projection."""
    G, I = 4 * Vz * V * V, 4 * V * V
    return 14 * G + 72 * N + 4 * I


def stage_algorithmic_bytes(N, V, Vz):
    """Per-projection algorithmic bytes of each stage (terms of the formula above;
    the fused XY kernel does the work of the X stage (2G) and the Y stage (2G))."""
    G, I = 4 * Vz * V * V, 4 * V * V
    return {"memset": 0, "pose_scatter": G + 36 * N, "blur_xy_fwd": 4 * G,
            "blurz_drc_fwd": G + 2 * I, "drc_blurz_bwd": 2 * G + 2 * I,
            "blur_xy_bwd": 5 * G, "gather_pose_bwd": G + 36 * N}


def stage_fma(V, Vz, K):
    """FP32 FMAs per projection of each stage's blur passes (K taps per output voxel and pass):
    the plane kernels run the X and the Y pass (their adjoints backward), the ray kernels the Z
    pass; the DRC recurrence adds ~6 per voxel.  Point stages: not FMA-bound (None)."""
    vox = Vz * V * V
    return {"memset": None, "pose_scatter": None, "blur_xy_fwd": 2 * K * vox,
            "blurz_drc_fwd": (K + 6) * vox, "drc_blurz_bwd": (K + 8) * vox,
            "blur_xy_bwd": 2 * K * vox, "gather_pose_bwd": None}


# DRAM bytes per whole-batch launch (dram__bytes_read.sum + dram__bytes_write.sum) of each stage's
# kernel(s) at workload A, from the committed `ncu --set full` capture NCU_CAPTURE (cold L2: ncu
# flushes the caches before every replay).  None = not captured for that workload.
NCU_CAPTURE = "profiles/r02_ncu_full_step.csv"
NCU_TRAFFIC_BYTES = {
    "A": {"pose_scatter": 6.18e6 + 0.0, "blur_xy_fwd": 7.47e6 + 14.14e6,
          "blurz_drc_fwd": 67.15e6 + 13.71e6, "drc_blurz_bwd": 71.36e6 + 11.37e6,
          "blur_xy_bwd": 77.26e6 + 4.60e6, "gather_pose_bwd": 23.12e6 + 0.0},
}
# The same from a capture WITHOUT cache flushes of the replayed CUDA graph (`ncu --cache-control
# none --graph-profiling node`, scripts/profile_graph.py; profiles/r02_ncu_steady_state_graph_replay
# .csv): DRAM read + write per HALF-batch launch with L2 as the previous kernels of the chain left
# it (ncu still serialises the two half-batch chains).  x 2 = per whole batch.
NCU_STEADY_CAPTURE = "profiles/r02_ncu_steady_state_graph_replay.csv"
NCU_STEADY_BYTES_HALF = {
    "A": {"pose_scatter": 1.41e6 + 0.13e6, "blur_xy_fwd": 3.76e6 + 1.73e6,
          "blurz_drc_fwd": 0.50e6 + 2.05e6, "drc_blurz_bwd": 35.64e6 + 4.32e6,
          "blur_xy_bwd": 37.31e6 + 0.0, "gather_pose_bwd": 8.72e6 + 0.0},
}


def make_cfg(w):
    from pytorch_unsup_pc_b200.config import default_cfg
    return default_cfg(vox_size=w["V"], pc_gauss_kernel_size=w["K"],
                       pose_predict_num_candidates=max(w.get("cands", 1), 1))


def synth_inputs(w, seed, P=None):
    """SURVEY.md 8(d) primary inputs, generated on the CPU with a fixed seed."""
    import torch
    g = torch.Generator().manual_seed(seed)
    P = P or w["P"]
    pts = (torch.rand(P, w["N"], 3, generator=g) - 0.5) * 0.9
    quat = torch.randn(P, 4, generator=g)
    scale = 0.2 + 0.8 * torch.rand(P, 1, generator=g)
    g5 = torch.Generator().manual_seed(5)
    Wp = torch.rand(P, w["V"], w["V"], 1, generator=g5)
    Wd = 0.1 * torch.rand(P, w["V"], w["V"], 1, generator=g5)
    return dict(points=pts, quat=quat, scale=scale, g_mask=Wp, g_depth=Wd)


def synth_masks(w, seed, BV=None):
    """Ground-truth masks of the loss legs: Bernoulli(0.5) G x G images."""
    import torch
    g = torch.Generator().manual_seed(seed + 7919)
    BV = BV or w["clouds"] * w["views"]
    return (torch.rand(BV, 1, w["G"], w["G"], generator=g) > 0.5).float()


# ----------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self, t0, t1):
        """Summarise the samples taken while the GPU was under load, t0 <= t <= t1."""
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, r in self.rows:
            if t < t0 or t > t1:
                continue
            try:
                sm.append(float(r[0]))
                smax = max(smax, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU legs (oracle)
def cpu_reference_rate(w, P_cpu, reps, warm, threads):
    """fwd+bwd projections/s of the oracle port on `threads` host threads."""
    import torch
    from oracle import closed_form as CF
    torch.set_num_threads(threads)
    cfg = make_cfg(w)
    kern = CF.smoothing_taps(cfg, w["sigma"])
    inp = synth_inputs(w, 2000, P=P_cpu)
    times = []
    for it in range(warm + reps):
        leaves = [inp[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
        t0 = time.perf_counter()
        out = CF.project(cfg, leaves[0], leaves[1], None, kern, leaves[2])
        torch.autograd.backward([out["proj"], out["proj_depth"]],
                                [inp["g_mask"].double(), inp["g_depth"].double()])
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    times.sort()
    return P_cpu / times[len(times) // 2], times


def cpu_fused_rate(w, clouds, reps, warm, threads):
    """projections/s of the oracle composition of the fused step (replication + projection +
    candidate-selection loss + autograd) on `clouds` clouds x views x candidates."""
    import torch
    from oracle import closed_form as CF
    from oracle import render_loss as ORL
    torch.set_num_threads(threads)
    cfg = make_cfg(w)
    kern = CF.smoothing_taps(cfg, w["sigma"])
    R = w["views"] * w["cands"]
    inp = synth_inputs(w, 2000, P=clouds * R)
    pts = inp["points"][::R].contiguous()
    masks = synth_masks(w, 2000, BV=clouds * w["views"])
    times = []
    for it in range(warm + reps):
        leaves = [t.clone().requires_grad_() for t in (pts, inp["quat"], inp["scale"])]
        t0 = time.perf_counter()
        loss, _, _ = ORL.project_candidates_loss(cfg, leaves[0], leaves[1], masks, w["cands"], kern,
                                                 leaves[2])
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    times.sort()
    return clouds * R / times[len(times) // 2]


def oracle_parity(cfg, kern, inp, got, idx):
    """The CUDA results `got` (mask, depth, g_points, g_quat, g_scale; whole batch, on the CPU)
    of the inputs `inp` against the oracle for the projections `idx`."""
    import torch
    from oracle import closed_form as CF
    sl = torch.tensor(idx)
    leaves = [inp[k][sl].clone().requires_grad_() for k in ("points", "quat", "scale")]
    out = CF.project(cfg, leaves[0], leaves[1], None, kern, leaves[2])
    torch.autograd.backward([out["proj"], out["proj_depth"]],
                            [inp["g_mask"][sl].double(), inp["g_depth"][sl].double()])

    def rel(a, b):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    fwd = {"mask": rel(got["mask"][sl], out["proj"].detach()),
           "depth": rel(got["depth"][sl], out["proj_depth"].detach())}
    grad = {"points": rel(got["g_points"][sl], leaves[0].grad),
            "quat": rel(got["g_quat"][sl], leaves[1].grad),
            "scale": rel(got["g_scale"][sl], leaves[2].grad)}
    return {"n": len(idx), "projections": list(idx), "against": "oracle/closed_form.py (fp64), same inputs",
            "fwd": fwd, "grad": grad, "max_fwd": max(fwd.values()), "max_grad": max(grad.values()),
            "tol_fwd": FWD_TOL, "tol_grad": GRAD_TOL,
            "tol_ok": max(fwd.values()) < FWD_TOL and max(grad.values()) < GRAD_TOL}


def oracle_fused_parity(cfg, kern, w, pts, quat, scale, masks, got, clouds):
    """The fused step's results against the oracle composition, for the clouds `clouds` (all of
    their views and candidates).  The loss of a view depends on that view only, so a subset of
    clouds is a closed sub-problem: weight_scale is rescaled by the subset's share of the views."""
    import torch
    from oracle import render_loss as ORL
    R, C, views = w["views"] * w["cands"], w["cands"], w["views"]
    cl = torch.tensor(clouds)
    pr = torch.cat([torch.arange(c * R, (c + 1) * R) for c in clouds])
    bv = torch.cat([torch.arange(c * views, (c + 1) * views) for c in clouds])
    BV_all = pts.shape[0] * views
    leaves = [pts[cl].clone().requires_grad_(), quat[pr].clone().requires_grad_(),
              scale[pr].clone().requires_grad_()]
    loss, min_loss, proj = ORL.project_candidates_loss(cfg, leaves[0], leaves[1], masks[bv], C, kern,
                                                       leaves[2], weight_scale=len(bv) / BV_all)
    loss.backward()

    def rel(a, b):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    fwd = {"projs": rel(got["mask"][pr], proj.detach())}
    grad = {"points": rel(got["g_points"][cl], leaves[0].grad),
            "quat": rel(got["g_quat"][pr], leaves[1].grad),
            "scale": rel(got["g_scale"][pr], leaves[2].grad)}
    argmin_ok = got["min_idx"][bv].tolist() == min_loss.tolist()
    return {"n": len(pr), "clouds": list(clouds), "against": "oracle/render_loss.py (replicas + "
            "closed_form + loss, fp64), same inputs", "fwd": fwd, "grad": grad, "argmin_equal": argmin_ok,
            "max_fwd": max(fwd.values()), "max_grad": max(grad.values()),
            "tol_fwd": FWD_TOL, "tol_grad": GRAD_TOL,
            "tol_ok": argmin_ok and max(fwd.values()) < FWD_TOL and max(grad.values()) < GRAD_TOL}


def run_reference(args, rank, world):
    """The reference arm: the path's CPU implementation on this host's cores.
    The reference is Python and is not present on the GPU box, so the timed
    code is the oracle port (same torch ops, same dtype flow; oracle/closed_form.py)."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    import torch
    from oracle import closed_form as CF
    torch.set_num_threads(threads)
    cfg = make_cfg(w)
    kern = CF.smoothing_taps(cfg, w["sigma"])
    # bounded sample: size the per-step batch so the whole run stays near 2 minutes
    _, t1 = cpu_reference_rate(w, 1, reps=1, warm=1, threads=threads)
    P_cpu = int(max(1, min(8, 120.0 / ((args.steps + args.warmup) * t1[0]))))
    inp = synth_inputs(w, 2000, P=P_cpu)

    def step():
        leaves = [inp[k].clone().requires_grad_() for k in ("points", "quat", "scale")]
        out = CF.project(cfg, leaves[0], leaves[1], None, kern, leaves[2])
        torch.autograd.backward([out["proj"], out["proj_depth"]],
                                [inp["g_mask"].double(), inp["g_depth"].double()])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = P_cpu * args.steps / dt
    sample = "%d projections/step of workload %s (N=%d, %d^3, K=%d)" % (
        P_cpu, args.workload, w["N"], w["V"], w["K"])
    # the other legs of the B200 arm, each on a small bounded sample (informational)
    extra = {}
    if args.workload == "A" and not args.main_only:
        extra["fused"] = {"value": cpu_fused_rate(w, 1, reps=1, warm=1, threads=threads), "unit": UNIT,
                          "sample": "1 cloud x %d candidates: replication + projection + candidate-"
                                    "selection loss + autograd (oracle/render_loss.py)" % w["cands"]}
        vb, _ = cpu_reference_rate(WORKLOADS["B"], 1, reps=1, warm=1, threads=threads)
        extra["workloads"] = {"B": {"value": vb, "unit": UNIT, "sample": "1 projection, 1 warm-up + 1"},
                              "C3": {"value": value, "unit": UNIT, "sample": "same per-projection work as A"},
                              "C5": {"value": value, "unit": UNIT, "sample": "same per-projection work as A "
                                     "(the reference's index_put_ scatter is deterministic on the CPU)"}}
    print(json.dumps(dict({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w["name"], "P_per_step": P_cpu, "N": w["N"], "V": w["V"],
                   "K": w["K"], "sigma": w["sigma"], "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }, **extra)))


# ----------------------------------------------------------------------------
class Env:
    """What every leg needs: the device, the library, the distributed fences and the timer."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import pytorch_unsup_pc_b200 as dpc
        from pytorch_unsup_pc_b200 import _lib
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch, self.dist, self.dpc = torch, dist, dpc
        self.lib = _lib.load()
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.current_stream(self.dev)

    def fence(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = self.torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(self, fn, steps, warmup, join=()):
        """CUDA events on the launching stream over exactly `steps` steps; `join` = side
        streams whose work (the last step's copies) must be inside the timed region."""
        torch = self.torch
        for i in range(warmup):
            fn(i)
        self.fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for i in range(steps):
            fn(warmup + i)
        for side in join:
            self.stream.wait_stream(side)
        e1.record(self.stream)
        self.fence()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def graph_of(self, fn, n, sptr):
        """`fn(i)` for i < n captured once (the C-ABI calls take the stream from `sptr`)."""
        torch = self.torch
        for i in range(n):
            fn(i)
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(self.dev)
        cap.wait_stream(self.stream)
        with torch.cuda.stream(cap):
            saved = sptr.value
            sptr.value = cap.cuda_stream
            with torch.cuda.graph(graph, stream=cap):
                for i in range(n):
                    fn(i)
            sptr.value = saved
        self.stream.wait_stream(cap)
        return graph


def measure_fma_peak(env):
    """FFMA2 rate of this GPU (scalar FMAs per second), all SMs busy at the occupancy of the
    blur kernels' inner loops: best of 5 launches of dpc_fma_rate_probe."""
    torch = env.torch
    blocks, iters = 148 * 8, 2000
    out = torch.empty(blocks * 256, dtype=torch.float32, device=env.dev)
    cnt = ctypes.c_double(0.0)
    sp = ctypes.c_void_p(env.stream.cuda_stream)
    best = 0.0
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        st = env.lib.dpc_fma_rate_probe(blocks, iters if it else 10, out.data_ptr(), ctypes.byref(cnt), sp)
        e1.record(env.stream)
        torch.cuda.synchronize(env.dev)
        if st != 0:
            return None
        if it:
            best = max(best, cnt.value / (e0.elapsed_time(e1) * 1e-3))
    return best


def measure_link(env, h2d_bytes, d2h_bytes, P, steps=60):
    """projections/s at which this box moves one step's bytes per step: one H2D and one D2H copy
    of the step's sizes per step on two streams, both directions at once, every rank at once."""
    torch = env.torch
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty_like(h_in, device=env.dev), torch.empty_like(h_out, device=env.dev)
    s1, s2 = torch.cuda.Stream(env.dev), torch.cuda.Stream(env.dev)

    def step(i):
        if i == 0:
            s1.wait_stream(env.stream)
            s2.wait_stream(env.stream)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    ms = env.timed(step, steps, 10, join=(s1, s2))
    return {"value": env.world * P * steps / (ms * 1e-3), "unit": UNIT, "us_per_step": ms / steps * 1e3,
            "how": "one %d-byte H2D and one %d-byte D2H copy per step on two streams, all %d rank(s) "
                   "at once, CUDA events, max over ranks" % (h2d_bytes, d2h_bytes, env.world)}


def bench_projection(env, key, full):
    """One workload: `value`, parity of the timed buffers, `e2e` (+ link ceiling), per-stage
    timings and roofline; `full` adds the informational legs of the headline workload."""
    import torch
    from pytorch_unsup_pc_b200 import _lib, ops
    args, dpc, lib, dev, world, rank = env.args, env.dpc, env.lib, env.dev, env.world, env.rank
    stream = env.stream
    w = WORKLOADS[key]
    cfg = make_cfg(w)
    P, N, V = w["P"], w["N"], w["V"]
    Vz = V
    deterministic = bool(w.get("deterministic"))
    mode = _lib.SCATTER_SORTED if deterministic else _lib.SCATTER_ATOMIC
    use_cells = not args.global_grid       # both scatter modes run plane-local and save the same state
    kern = dpc.smoothing_kernel(cfg, w["sigma"])
    taps = ops.host_taps(kern)
    params = ops.make_params(cfg, P, N, flip_y=True)
    steps = args.steps

    # ---- inputs: N_INPUT_SETS seeded sets, pinned on the host, resident on the device ----
    host = [synth_inputs(w, 1000 + 17 * rank + i) for i in range(N_INPUT_SETS)]
    for h in host:
        for k in h:
            h[k] = h[k].contiguous().pin_memory()
    devin = [{k: v.to(dev) for k, v in h.items()} for h in host]
    f32 = dict(dtype=torch.float32, device=dev)

    def make_buffers():
        return dict(tr_pc=torch.empty(P, N, 3, **f32), grid=torch.empty(P, Vz, V, V, **f32),
                    bits=torch.empty(P, Vz, V, V // 32, dtype=torch.int32, device=dev),
                    mask=torch.empty(P, V, V, **f32), depth=torch.empty(P, V, V, **f32),
                    g_grid=torch.empty(P, Vz, V, V, **f32), g_points=torch.empty(P, N, 3, **f32),
                    g_quat=torch.empty(P, 4, **f32), g_scale=torch.empty(P, **f32),
                    cells=torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8,
                                      device=dev) if use_cells else None,
                    ws=torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8,
                                   device=dev))
    buf = make_buffers()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    P_ = ops._ptr
    tap_args = ops._tap_args(taps)

    def step_abi(i, buf=buf):
        d = devin[i % N_INPUT_SETS]
        ws = buf["ws"]
        st = lib.dpc_project_fwd(ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None,
                                 P_(d["scale"]), *tap_args, mode, P_(buf["tr_pc"]),
                                 P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]), P_(buf["mask"]),
                                 P_(buf["depth"]), None, None, P_(ws), ws.numel(), sptr)
        _lib.check(st, "project_fwd")
        st = lib.dpc_project_bwd(ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None,
                                 P_(d["scale"]), *tap_args, P_(buf["grid"]), P_(buf["bits"]),
                                 P_(buf["cells"]), P_(d["g_mask"]), P_(d["g_depth"]), None, None, None,
                                 P_(buf["g_grid"]), P_(buf["g_points"]), P_(buf["g_quat"]), None,
                                 None, P_(buf["g_scale"]), P_(ws), ws.numel(), sptr)
        _lib.check(st, "project_bwd")

    sampler = ClockSampler(env.local_rank) if (rank == 0 and full) else None
    if sampler:
        sampler.start()
        time.sleep(0.5)
    # ---- (1) value: device-resident inputs, C-ABI calls ----
    t_load0 = time.time()
    graph = None
    if args.graph:
        # the N_INPUT_SETS steps captured once into a CUDA graph and replayed
        graph = env.graph_of(step_abi, N_INPUT_SETS, sptr)
        reps = (steps + N_INPUT_SETS - 1) // N_INPUT_SETS

        def replay(i):
            if i % N_INPUT_SETS == 0:
                graph.replay()
        ms_abi = env.timed(replay, reps * N_INPUT_SETS, N_INPUT_SETS * 2) * steps / (reps * N_INPUT_SETS)
    else:
        ms_abi = env.timed(step_abi, steps, max(args.warmup, 3))
    # ---- (1a) parity of what was just timed: the buffers hold the LAST timed step's results
    # (input set N_INPUT_SETS - 1 under graph replay) -> a sample of projections from both
    # half-batches against the CPU oracle on the same inputs
    last = (N_INPUT_SETS - 1) if args.graph else (max(args.warmup, 3) + steps - 1) % N_INPUT_SETS
    parity = None
    if rank == 0 and not args.no_parity_check:
        got = {k: buf[k].cpu() for k in ("mask", "depth", "g_points", "g_quat", "g_scale")}
        k_s = 2 if V >= 128 else 4
        idx = sorted({0, P // 2 - 1, P // 2, P - 1})[:k_s] if k_s == 4 else [0, P - 1]
        parity = oracle_parity(cfg, kern, host[last], got, idx)
        parity["input_set"] = last
        if deterministic:
            # configs[4]: bit-exact repeatability -- the same step again into fresh buffers
            buf2 = make_buffers()
            step_abi(last, buf2)
            torch.cuda.synchronize(dev)
            parity["bit_exact_repeat"] = all(torch.equal(buf[k], buf2[k]) for k in
                                             ("mask", "depth", "g_points", "g_quat", "g_scale", "tr_pc"))
            parity["tol_ok"] = parity["tol_ok"] and parity["bit_exact_repeat"]
            del buf2
    # ---- (1b) informational: the same steps fed as TWO independent streams of batches (a second
    # set of buffers and a second graph on a second stream, replayed alternately): what the GPU
    # sustains when a step's kernel tails are filled by another step's kernels.  `value` stays the
    # one-stream figure.
    ms_two = None
    if full and args.graph and use_cells:
        buf_b = make_buffers()
        lane_streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        graph2 = env.graph_of(lambda i: step_abi(i, buf_b), N_INPUT_SETS, sptr)
        lanes = (graph, graph2)

        def replay2(i):
            if i % N_INPUT_SETS == 0:
                l = (i // N_INPUT_SETS) % 2
                lane_streams[l].wait_stream(stream)
                with torch.cuda.stream(lane_streams[l]):
                    lanes[l].replay()
        reps2 = reps + (reps % 2)
        ms_two = env.timed(replay2, reps2 * N_INPUT_SETS, N_INPUT_SETS * 4, join=lane_streams) * steps / (
            reps2 * N_INPUT_SETS)
        del buf_b, graph2
    clocks = None
    if sampler:
        # the timed region can be shorter than nvidia-smi's sampling period: keep
        # the same step running until the sampler has seen >= 1.5 s under load
        k = 0
        while time.time() < t_load0 + 1.5:
            step_abi(k)
            k += 1
            if k % 50 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
        clocks = sampler.stop(t_load0 + 0.2, time.time())

    # ---- (2) e2e: public Python API, pinned host inputs in, results out, every step ----
    dpc.set_outputs(voxels=False, drc_probs=False)
    dpc.point_cloud._options["plane_local"] = not args.global_grid
    dpc.set_deterministic(deterministic)
    E2E_LANES = int(os.environ.get("DPC_E2E_LANES", "2"))
    g_steps = E2E_GRAPH_STEPS if full else 2 * N_INPUT_SETS

    def make_e2e(project, host_sets, out_shapes, nbytes_in, nbytes_out):
        """lane -> step function: this step's inputs pinned host memory -> device (copy stream,
        overlapped with the previous step's kernels); results and gradients device -> pinned host
        memory (readback stream, overlapped with the next step's kernels).  Every lane has its
        own pipeline and its own pinned result buffers."""
        def make_step(lane, pipe):
            out_h = {k: torch.empty(*shp).pin_memory() for k, shp in out_shapes.items()}

            def step(i):
                h = host_sets[i % N_INPUT_SETS]
                din = pipe.upload({"points": h["points"], "quat": h["quat"], "scale": h["scale"]})
                pts = din["points"].detach().requires_grad_()
                quat = din["quat"].detach().requires_grad_()
                scale = din["scale"].detach().requires_grad_()
                d = devin[i % N_INPUT_SETS]
                out = project(cfg, pts, quat, None, None, kern, scaling_factor=scale)
                gp, gq, gs = torch.autograd.grad([out["proj"], out["proj_depth"]], [pts, quat, scale],
                                                 [d["g_mask"], d["g_depth"]])
                pipe.download({"mask": out["proj"], "depth": out["proj_depth"], "g_points": gp,
                               "g_quat": gq, "g_scale": gs}, out_h)
                assert pipe.h2d_bytes == nbytes_in and pipe.d2h_bytes == nbytes_out
            return step
        return make_step

    def measure_e2e(make_step):
        if args.graph:
            # g_steps consecutive e2e steps (each with its own H2D and D2H copies) captured
            # once per lane with the public helpers and replayed: the eager loop is bound by the
            # host (~0.25 ms of Python / autograd-engine work per ~0.15 ms step); two lanes
            # replayed in turn keep the copy pipeline from draining at every replay boundary
            ag = dpc.AlternatingGraphs(make_step, g_steps, dev, lanes=E2E_LANES, warmup=2)
            # at least 8 replays (the first replay of every lane ramps its copy pipeline up), a
            # multiple of the lane count; the time is scaled to `steps` below
            reps = max(8, (steps + g_steps - 1) // g_steps)
            reps += reps % E2E_LANES and (E2E_LANES - reps % E2E_LANES)

            def replay_e2e(i):
                if i % g_steps == 0:
                    ag.replay()
            ms = env.timed(replay_e2e, reps * g_steps, 2 * E2E_LANES * g_steps,
                           join=ag.streams) * steps / (reps * g_steps)
            mode_s = "AlternatingGraphs(%d lanes x %d steps per CUDA graph)" % (E2E_LANES, g_steps)
            for p_ in ag.pipes:
                p_.drain()
        else:
            pipe = dpc.HostPipeline(dev, depth=3)
            ms = env.timed(make_step(0, pipe), steps, max(args.warmup, 3), join=(pipe.h2d, pipe.d2h))
            mode_s = "eager"
            pipe.drain()
        return ms, mode_s

    out_shapes = dict(mask=(P, V, V, 1), depth=(P, V, V, 1), g_points=(P, N, 3), g_quat=(P, 4),
                      g_scale=(P, 1))
    h2d = sum(host[0][k].numel() * 4 for k in ("points", "quat", "scale"))
    d2h = sum(4 * int(torch.Size(shp).numel()) for shp in out_shapes.values())
    ms_e2e, e2e_mode = measure_e2e(make_e2e(dpc.pointcloud_project_fast, host, out_shapes, h2d, d2h))
    link = measure_link(env, h2d, d2h, P)

    # ---- (2b) the same step through the replica-aware API (next row f2): the host holds the
    # UN-replicated clouds (P / replicas of them), the kernels read cloud b // replicas, and
    # the cloud gradient comes back summed over the replicas -- fewer point bytes each way
    R = w["views"] * w["cands"]
    e2e_rep = None
    if full and R > 1 and P % R == 0:
        B = P // R
        host_rep = [{"points": h["points"][::R].contiguous().pin_memory(), "quat": h["quat"],
                     "scale": h["scale"]} for h in host]
        shapes_rep = dict(out_shapes, g_points=(B, N, 3))
        h2d_rep = sum(host_rep[0][k].numel() * 4 for k in ("points", "quat", "scale"))
        d2h_rep = sum(4 * int(torch.Size(shp).numel()) for shp in shapes_rep.values())
        ms_rep, mode_rep = measure_e2e(make_e2e(dpc.pointcloud_project_replicated, host_rep,
                                                shapes_rep, h2d_rep, d2h_rep))
        e2e_rep = {"value": world * P * steps / (ms_rep * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": h2d_rep, "d2h_bytes_per_step": d2h_rep,
                   "ms_per_step": ms_rep / steps,
                   "api": "pytorch_unsup_pc_b200.pointcloud_project_replicated: %d clouds x %d "
                          "replicas per step, cloud gradient summed over the replicas in the "
                          "kernels" % (B, R), "mode": mode_rep}

    # ---- (2c) the fused step: renderer + candidate-selection loss (f2 -> path -> f1) ----
    fused = None
    if w["cands"] > 1 and not deterministic and not args.global_grid and not args.no_fused:
        fused = bench_fused(env, key, w, cfg, kern, taps, host, measure_e2e, g_steps, full)

    # ---- (3) per-stage CUDA-event timings for the roofline ----
    stage_ms = (ctypes.c_float * len(_lib.PROFILE_STAGES))()
    d = devin[0]
    ws = buf["ws"]
    st = lib.dpc_project_profile(
        ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]), *tap_args,
        mode, P_(buf["tr_pc"]), P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]),
        P_(buf["mask"]), P_(buf["depth"]), P_(d["g_mask"]), P_(d["g_depth"]), P_(buf["g_grid"]), P_(buf["g_points"]),
        P_(buf["g_quat"]), None, None, P_(buf["g_scale"]), P_(ws), ws.numel(), sptr,
        min(max(steps, 10), 50), stage_ms)
    _lib.check(st, "project_profile")
    env.fence()
    dpc.set_deterministic(False)

    stages = dict(zip(_lib.PROFILE_STAGES, [float(x) for x in stage_ms]))
    n_chunks = lib.dpc_project_chunks(ctypes.byref(params))
    sbytes = stage_algorithmic_bytes(N, V, Vz)
    sfma = stage_fma(V, Vz, w["K"])
    top = max((k for k in stages if k != "memset"), key=lambda k: stages[k])
    achieved = sbytes[top] * P / (stages[top] * 1e-3) / 1e9
    value = world * P * steps / (ms_abi * 1e-3)
    e2e_value = world * P * steps / (ms_e2e * 1e-3)
    peak, peak_src = env.hbm_peak
    step_frac = (value / world) * algorithmic_bytes(N, V, Vz) / 1e9 / peak
    traffic = None if (args.global_grid or deterministic) else NCU_TRAFFIC_BYTES.get(key, {}).get(top)
    steady = None if (args.global_grid or deterministic) else NCU_STEADY_BYTES_HALF.get(key, {}).get(top)
    fma_peak = env.fma_peak
    fp32 = None
    if fma_peak:
        per_stage = {k: (sfma[k] * P / (stages[k] * 1e-3) / fma_peak) for k in stages
                     if sfma.get(k) and stages[k] > 0}
        fp32 = {"kernel_frac": per_stage.get(top), "per_stage": per_stage,
                "step_frac": sum(v for v in sfma.values() if v) * (value / world) / fma_peak,
                "peak_tfma_s": fma_peak / 1e12,
                "peak_source": "dpc_fma_rate_probe on this GPU in this run (FFMA2, 8 independent packed "
                               "accumulators per thread, 148 x 8 CTAs of 256 threads, best of 5)",
                "fma_per_projection": {k: v for k, v in sfma.items() if v}}

    # ---- (4) CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline and rank == 0:
        threads = os.cpu_count() or 1
        P_cpu = 1 if V >= 128 else (8 if full else 2)
        v, times = cpu_reference_rate(w, P_cpu, reps=3 if full else 1, warm=1, threads=threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d projections of the same workload, fwd+bwd, 1 warm-up + median of %d "
                         "(oracle/closed_form.py: the reference's torch ops in fp64)" % (
                             P_cpu, 3 if full else 1)}
    rec = {
        "value": value, "unit": UNIT, "ms_per_step": ms_abi / steps, "steps": steps,
        "config": {"workload": w["name"], "P_per_gpu": P, "N": N, "V": V, "Vz": Vz, "K": w["K"],
                   "sigma": w["sigma"],
                   "scatter": ("deterministic sort-then-segment, raw grid in global memory"
                               if deterministic and args.global_grid else
                               "deterministic sort-then-segment, plane-local (records sorted by grid "
                               "row, every plane row summed in a fixed order in shared memory)"
                               if deterministic else
                               "global grid, atomic" if args.global_grid else "plane-local (shared memory)"),
                   "l2": "no flush: %d input sets rotate and each step's grid + gradient-grid "
                         "working set (%d MiB) exceeds the 126 MB L2" % (
                             N_INPUT_SETS, 2 * P * Vz * V * V * 4 >> 20),
                   "parallelism": "projections sharded across ranks, no collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / steps,
                "api": "pytorch_unsup_pc_b200.pointcloud_project_fast + torch.autograd.grad, "
                       "host copies through pytorch_unsup_pc_b200.HostPipeline (3 streams: this "
                       "step's H2D / kernels / D2H overlap the neighbouring steps'), steps "
                       "captured and replayed with pytorch_unsup_pc_b200.AlternatingGraphs",
                "mode": e2e_mode, "link_ceiling": link,
                "frac_of_link_ceiling": e2e_value / link["value"]},
        "e2e_replica_aware": e2e_rep,
        "fused": fused,
        "parity_check": parity,
        "throughput_two_streams": None if ms_two is None else {
            "value": world * P * steps / (ms_two * 1e-3), "unit": UNIT,
            "ms_per_step": ms_two / steps,
            "note": "informational: the same device-resident steps issued as two independent "
                    "streams of batches (two graphs, two streams, replayed alternately); `value` "
                    "is the one-stream figure"},
        # kernels per chunk: pose_bin (pose_cells + bin_points above 16384 points; pose_scatter on
        # the global-grid path), blur_xy, blurz_drc_fwd | drc_blurz_bwd, blur_xy, gather_pose_bwd
        # -- times the chunks the batch is split into (whole job: every rank launches its own);
        # the deterministic mode: records + sort kernels instead of pose_bin (7 per chunk; 8 with
        # the segment kernel of the global-grid variant)
        "gpu_launches": ((8 if args.global_grid else 7) if deterministic else 6 if args.global_grid else
                         lib.dpc_project_kernels_per_chunk(ctypes.byref(params)))
                        * n_chunks * steps * world,
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "frac_label": "contract bytes (SURVEY 8d stage term x projections per launch) / "
                                   "kernel time / HBM peak; may pass 1: the design never moves the raw grid",
                     "traffic": traffic,
                     "traffic_source": NCU_CAPTURE + " (ncu --set full, dram read + write of one "
                                       "whole-batch launch, cold L2)",
                     "note": "kernel_ms and the byte counts are per whole-batch launch "
                             "(dpc_project_profile runs one chunk); the timed step runs %d "
                             "half-batch launches of every kernel on two streams" % n_chunks,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": sbytes[top] * P,
                     "kernel_ms": stages[top],
                     # the same kernel against the bytes it really moves (ncu) -- far below the
                     # HBM peak: its ceilings are shared-memory bandwidth and FP32 FMA issue
                     "traffic_gbs": None if traffic is None else traffic / (stages[top] * 1e-3) / 1e9,
                     "dram_frac_cold": None if traffic is None else traffic / (stages[top] * 1e-3) / 1e9 / peak,
                     # steady state: no cache flush, replayed graph (2 x the half-batch launch)
                     "dram_frac_steady": None if steady is None else 2 * steady / (stages[top] * 1e-3) / 1e9 / peak,
                     "traffic_steady": None if steady is None else 2 * steady,
                     "traffic_steady_source": NCU_STEADY_CAPTURE,
                     "fp32_frac": None if fp32 is None else fp32["kernel_frac"],
                     "fp32": fp32,
                     "limiter": "shared-memory bandwidth / FP32 FMA issue, not HBM"},
        "roofline_step": {"algorithmic_bytes_per_projection": algorithmic_bytes(N, V, Vz),
                          "frac": step_frac, "peak": peak, "unit": "GB/s"},
        "stage_ms": stages,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    del buf, devin, graph
    torch.cuda.empty_cache()
    return rec


def bench_trajectory(env, key="A"):
    """`value` along the training trajectory of the reference: sigma_rel falls 3.0 -> 0.2
    (model_pc_to.py:59-63; the kernels run with the tap radius that holds all but 1e-7 of the taps)
    while the point-dropout keep probability rises 0.07 -> 1.0 (:68-87; N_eff = 560 -> 8000).
    Device-resident C-ABI steps replayed from a CUDA graph, this rank only (no fence)."""
    import torch
    from pytorch_unsup_pc_b200 import _lib, ops
    dpc, lib, dev, stream = env.dpc, env.lib, env.dev, env.stream
    out = []
    for keep in (0.07, 0.5, 1.0):
        for sigma in (3.0, 1.0, 0.5, 0.2):
            w = dict(WORKLOADS[key], sigma=sigma, N=max(1, int(WORKLOADS[key]["N"] * keep)))
            cfg = make_cfg(w)
            P, N, V = w["P"], w["N"], w["V"]
            taps = ops.host_taps(dpc.smoothing_kernel(cfg, sigma))
            params = ops.make_params(cfg, P, N, flip_y=True)
            d = {k: v.to(dev) for k, v in synth_inputs(w, 1000).items()}
            f32 = dict(dtype=torch.float32, device=dev)
            buf = dict(tr_pc=torch.empty(P, N, 3, **f32), grid=torch.empty(P, V, V, V, **f32),
                       bits=torch.empty(P, V, V, V // 32, dtype=torch.int32, device=dev),
                       mask=torch.empty(P, V, V, **f32), depth=torch.empty(P, V, V, **f32),
                       g_grid=torch.empty(P, V, V, V, **f32), g_points=torch.empty(P, N, 3, **f32),
                       g_quat=torch.empty(P, 4, **f32), g_scale=torch.empty(P, **f32),
                       cells=torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev))
            ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
            sptr = ctypes.c_void_p(stream.cuda_stream)
            P_, ta = ops._ptr, ops._tap_args(taps)

            def step(i):
                _lib.check(lib.dpc_project_fwd(
                    ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]), *ta,
                    _lib.SCATTER_ATOMIC, P_(buf["tr_pc"]), P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]),
                    P_(buf["mask"]), P_(buf["depth"]), None, None, P_(ws), ws.numel(), sptr), "project_fwd")
                _lib.check(lib.dpc_project_bwd(
                    ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]), *ta,
                    P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]), P_(d["g_mask"]), P_(d["g_depth"]), None,
                    None, None, P_(buf["g_grid"]), P_(buf["g_points"]), P_(buf["g_quat"]), None, None,
                    P_(buf["g_scale"]), P_(ws), ws.numel(), sptr), "project_bwd")
            g = env.graph_of(step, 3, sptr)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(12):
                g.replay()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) / 36 * 1e3
            out.append({"sigma": sigma, "keep_prob": keep, "N_eff": N,
                        "tap_radius": lib.dpc_tap_radius(taps[0].data_ptr(), taps[0].numel()),
                        "us_per_step": us, "value_per_gpu": P / us * 1e6})
            del buf, ws, g, d
    torch.cuda.empty_cache()
    return {"unit": UNIT, "what": "device-resident fwd+bwd steps of workload %s shapes, one GPU (rank 0), "
            "36 graph-replayed steps per point" % key, "points": out}


def bench_fused(env, key, w, cfg, kern, taps, host, measure_e2e, g_steps, full):
    """The renderer + candidate-selection loss as one step (project_candidates_loss over
    dpc_render_loss_fwd / _bwd): clouds [B,N,3], poses [P,4], scales, ground-truth masks
    [BV,1,G,G] in; loss and cloud / pose / scale gradients out."""
    import torch
    from pytorch_unsup_pc_b200 import _lib, ops
    args, dpc, lib, dev, world, rank = env.args, env.dpc, env.lib, env.dev, env.world, env.rank
    stream = env.stream
    P, N, V, C, G = w["P"], w["N"], w["V"], w["cands"], w["G"]
    Vz = V
    R = w["views"] * C
    B, BV = P // R, P // C
    steps = args.steps
    params = ops.make_params(cfg, P, N, flip_y=True)
    params.outputs = 0
    f32 = dict(dtype=torch.float32, device=dev)
    hostf = [{"points": h["points"][::R].contiguous().pin_memory(), "quat": h["quat"],
              "scale": h["scale"], "masks": synth_masks(w, 1000 + 17 * rank + i).pin_memory()}
             for i, h in enumerate(host)]
    devf = [{k: v.to(dev) for k, v in h.items()} for h in hostf]
    slots = lib.dpc_render_loss_slots(ctypes.byref(params), C, 1, _lib.SCATTER_ATOMIC)
    buf = dict(grid=torch.empty(P, Vz, V, V, **f32),
               bits=torch.empty(P, Vz, V, V // 32, dtype=torch.int32, device=dev),
               cells=torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev),
               mask=torch.empty(P, V, V, **f32), all_loss=torch.empty(BV, C, **f32),
               min_idx=torch.empty(BV, dtype=torch.int64, device=dev),
               view_loss=torch.empty(BV, **f32), loss=torch.empty(1, **f32),
               winners=torch.empty(BV, dtype=torch.int32, device=dev), kcoef=torch.empty(BV, **f32),
               g_grid=torch.empty(slots, Vz, V, V, **f32), g_rep=torch.empty(slots, N, 3, **f32),
               g_mask=torch.empty(P, V, V, **f32) if slots == P else None,
               g_points=torch.empty(B, N, 3, **f32), g_quat=torch.empty(P, 4, **f32),
               g_scale=torch.empty(P, **f32),
               ws=torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev))
    sptr = ctypes.c_void_p(stream.cuda_stream)
    P_ = ops._ptr
    tap_args = ops._tap_args(taps)
    ws = buf["ws"]

    def step_abi(i):
        d = devf[i % N_INPUT_SETS]
        st = lib.dpc_render_loss_fwd(
            ctypes.byref(params), R, N, None, P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]),
            *tap_args, _lib.SCATTER_ATOMIC, C, G, P_(d["masks"]), None, ctypes.c_float(1.0),
            P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]), P_(buf["mask"]), P_(buf["all_loss"]),
            P_(buf["min_idx"]), P_(buf["view_loss"]), P_(buf["loss"]), P_(buf["winners"]),
            P_(buf["kcoef"]), P_(ws), ws.numel(), sptr)
        _lib.check(st, "render_loss_fwd")
        st = lib.dpc_render_loss_bwd(
            ctypes.byref(params), R, N, None, P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]),
            *tap_args, _lib.SCATTER_ATOMIC, C, G, P_(d["masks"]), None, ctypes.c_float(1.0),
            P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]), P_(buf["mask"]), P_(buf["min_idx"]),
            P_(buf["winners"]), P_(buf["kcoef"]), None, P_(buf["g_grid"]), P_(buf["g_rep"]), None,
            P_(buf["g_mask"]), P_(buf["g_points"]), P_(buf["g_quat"]), None, None, P_(buf["g_scale"]),
            P_(ws), ws.numel(), sptr)
        _lib.check(st, "render_loss_bwd")

    if args.graph:
        graph = env.graph_of(step_abi, N_INPUT_SETS, sptr)
        reps = (steps + N_INPUT_SETS - 1) // N_INPUT_SETS

        def replay(i):
            if i % N_INPUT_SETS == 0:
                graph.replay()
        ms = env.timed(replay, reps * N_INPUT_SETS, N_INPUT_SETS * 2) * steps / (reps * N_INPUT_SETS)
        last = N_INPUT_SETS - 1
    else:
        ms = env.timed(step_abi, steps, max(args.warmup, 3))
        last = (max(args.warmup, 3) + steps - 1) % N_INPUT_SETS
    parity = None
    if rank == 0 and not args.no_parity_check:
        got = {k: buf[k].cpu() for k in ("mask", "g_points", "g_quat", "g_scale", "min_idx")}
        h = hostf[last]
        parity = oracle_fused_parity(cfg, kern, w, h["points"], h["quat"], h["scale"], h["masks"], got,
                                     [0, B - 1])
        parity["input_set"] = last
        parity["loss"] = float(buf["loss"].item())
    # end to end through the public API
    out_shapes = dict(loss=(1,), min_loss=(BV,), g_points=(B, N, 3), g_quat=(P, 4), g_scale=(P, 1))
    h2d = sum(hostf[0][k].numel() * 4 for k in ("points", "quat", "scale", "masks"))
    d2h = 4 + 8 * BV + 4 * (B * N * 3 + P * 4 + P)

    def make_step(lane, pipe):
        out_h = {"loss": torch.empty(1).pin_memory(), "min_loss": torch.empty(BV, dtype=torch.int64).pin_memory(),
                 "g_points": torch.empty(B, N, 3).pin_memory(), "g_quat": torch.empty(P, 4).pin_memory(),
                 "g_scale": torch.empty(P, 1).pin_memory()}

        def step(i):
            h = hostf[i % N_INPUT_SETS]
            din = pipe.upload(h)
            pts = din["points"].detach().requires_grad_()
            quat = din["quat"].detach().requires_grad_()
            scale = din["scale"].detach().requires_grad_()
            out = dpc.project_candidates_loss(cfg, pts, quat, None, din["masks"], kern,
                                              scaling_factor=scale)
            gp, gq, gs = torch.autograd.grad(out["loss"], [pts, quat, scale])
            pipe.download({"loss": out["loss"].reshape(1), "min_loss": out["min_loss"], "g_points": gp,
                           "g_quat": gq, "g_scale": gs}, out_h)
            assert pipe.h2d_bytes == h2d and pipe.d2h_bytes == d2h, (pipe.h2d_bytes, h2d, pipe.d2h_bytes, d2h)
        return step
    ms_e2e, mode_s = measure_e2e(make_step)
    link = measure_link(env, h2d, d2h, P)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and rank == 0 and full:
        threads = os.cpu_count() or 1
        cpu = {"value": cpu_fused_rate(w, 2, reps=1, warm=1, threads=threads), "unit": UNIT,
               "cores": threads, "kind": "port",
               "sample": "2 clouds x %d replicas: replication + projection + candidate-selection loss "
                         "+ autograd (oracle/render_loss.py), 1 warm-up + 1" % R}
    value = world * P * steps / (ms * 1e-3)
    e2e_value = world * P * steps / (ms_e2e * 1e-3)
    rec = {"value": value, "unit": UNIT, "ms_per_step": ms / steps,
           "what": "replica-aware projection of %d clouds x %d views x %d candidates + candidate-"
                   "selection loss against %d^2 ground-truth masks, forward + backward; the backward "
                   "chain runs over the %d winning projections (the others' gradients are exactly "
                   "zero) with dL/dmask built inside the ray kernel" % (B, w["views"], C, G, slots),
           "api_value": "dpc_render_loss_fwd + dpc_render_loss_bwd (C ABI), device-resident inputs",
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / steps, "mode": mode_s,
                   "api": "pytorch_unsup_pc_b200.project_candidates_loss + torch.autograd.grad; clouds, "
                          "poses, scales and ground-truth masks from pinned host memory in, loss / argmin "
                          "/ cloud, pose and scale gradients out, every step",
                   "link_ceiling": link, "frac_of_link_ceiling": e2e_value / link["value"],
                   "frac_of_value": e2e_value / value},
           "parity_check": parity, "cpu_baseline": cpu}
    del buf, devf
    torch.cuda.empty_cache()
    return rec


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import pytorch_unsup_pc_b200 as dpc

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); "
                           "use --impl reference for the CPU arm")
    # one process per GPU: keep this rank (and the pinned staging buffers it allocates) on the
    # GPU's own NUMA node
    numa_cpus = dpc.bind_to_device_numa(local_rank)
    torch.cuda.set_device(local_rank)
    env = Env(args, rank, world, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=env.dev)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        env.hbm_peak = (float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)")
    else:
        env.hbm_peak = (6650.0, "fallback (B200_PROFILING.md)")
    env.fma_peak = measure_fma_peak(env)

    main = bench_projection(env, args.workload, full=True)
    if not args.main_only:
        main["trajectory"] = bench_trajectory(env, args.workload if args.workload in ("A", "C3") else "A")
    subs = {}

    def sub_record(fn, *a, **kw):
        # a sub-record that fails must not cost the headline line: at one GPU the error is
        # recorded in its place.  Under torchrun a rank that left a sub-record early would be
        # out of step with the others' collectives, so there the error propagates.
        if world > 1:
            return fn(*a, **kw)
        try:
            return fn(*a, **kw)
        except Exception as e:              # noqa: BLE001 (reported, not swallowed)
            sys.stderr.write("bench.py: sub-record failed: %r\n" % (e,))
            return {"error": "%s: %s" % (type(e).__name__, e)}

    if args.workload == "A" and not args.main_only:
        for key in ("C3", "B", "C5"):
            subs[key] = sub_record(bench_projection, env, key, full=False)
        if not args.no_train:
            try:
                from pytorch_unsup_pc_b200 import train_step
            except ImportError:
                train_step = None
            if train_step is not None:
                subs["train3"] = sub_record(train_step.bench, env, args)
    try:
        env.fence()
    except Exception:                       # noqa: BLE001
        if world > 1 or not any("error" in r for r in subs.values()):
            raise                           # only a failed sub-record may leave the device unusable
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    main["config"]["host_binding"] = (("rank bound to the %d CPUs local to its GPU (NVML affinity)"
                                       % len(numa_cpus)) if numa_cpus else "none")
    line = {"metric": METRIC, "value": main.pop("value"), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main.pop("ms_per_step"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    main.pop("unit"), main.pop("steps")
    line.update(main)
    if subs:
        line["workloads"] = subs
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="A", choices=sorted(WORKLOADS))
    ap.add_argument("--main-only", action="store_true",
                    help="only the headline workload (no `workloads` sub-records)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--global-grid", action="store_true",
                    help="A/B: keep the raw grid in global memory (memset + atomic scatter, grid "
                         "gather) instead of the default plane-local path")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="issue the C-ABI calls from the host every step instead of replaying "
                         "them from a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
