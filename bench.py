#!/usr/bin/env python
"""bench.py -- fwd+bwd point-cloud projections/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload A|B]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port on host cores)

A step = one forward + backward pass of the projection path over one batch of
synthetic projections (workload A: P = 16 x 4 pose candidates = 64 clouds of
8000 points -> 64^3 grid -> 64^2 mask + depth, K=21, sigma=3; BASELINE.json
configs[1]).  One process per GPU, projections sharded by (batch x candidate)
with no data-path collective (weak scaling: P per GPU is fixed).

`value`   device-resident inputs, the two C-ABI calls dpc_project_fwd /
          dpc_project_bwd per step, timed with CUDA events over exactly K steps
          (barrier + synchronize on both sides, max over ranks).
`e2e`     the same metric through the public Python API
          (pointcloud_project_fast + autograd) with pinned HOST inputs copied
          in and the results (mask, depth, gradients) copied out every step.
`roofline`      the longest kernel of the step, from per-stage CUDA-event
                timings taken in this run (dpc_project_profile).
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
              P=64, N=8000, V=64, K=21, sigma=3.0),
    # BASELINE.json configs[3]: paper scale
    "B": dict(name="paper scale: batch 32 x 4 candidates, 16000 pts -> 128^3 -> 128^2, K=21 "
                   "sigma=3.0, fwd+bwd",
              P=128, N=16000, V=128, K=21, sigma=3.0),
    # BASELINE.json configs[2], the projection part per GPU: batch 16 x 4 views x 4 candidates
    "C3": dict(name="chair_unsupervised train-step shapes per GPU: batch 16 x 4 views x 4 candidates, "
                    "8000 pts -> 64^3 -> 64^2, K=21 sigma=3.0, fwd+bwd",
               P=256, N=8000, V=64, K=21, sigma=3.0),
}
N_INPUT_SETS = 3          # distinct input sets rotated between steps
E2E_GRAPH_STEPS = int(os.environ.get("DPC_E2E_GRAPH_STEPS", "24"))   # e2e steps captured per CUDA graph (a multiple of N_INPUT_SETS)


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


# DRAM bytes per whole-batch launch (dram__bytes_read.sum + dram__bytes_write.sum) of each stage's
# kernel(s) at workload A, from the committed `ncu --set full` capture NCU_CAPTURE (cold L2: ncu
# flushes the caches before every replay).  None = not captured for that workload.
NCU_CAPTURE = "profiles/r01_ncu_full_step_pdl.csv"
NCU_TRAFFIC_BYTES = {
    "A": {"pose_scatter": 6.18e6 + 0.0, "blur_xy_fwd": 7.47e6 + 11.24e6,
          "blurz_drc_fwd": 67.15e6 + 12.97e6, "drc_blurz_bwd": 71.36e6 + 23.46e6,
          "blur_xy_bwd": 77.02e6 + 3.43e6, "gather_pose_bwd": 23.12e6 + 0.0},
}


def make_cfg(w):
    from pytorch_unsup_pc_b200.config import default_cfg
    return default_cfg(vox_size=w["V"], pc_gauss_kernel_size=w["K"])


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


# ----------------------------------------------------------------------------
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
    print(json.dumps({
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
    }))


# ----------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import pytorch_unsup_pc_b200 as dpc
    from pytorch_unsup_pc_b200 import _lib, ops

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); "
                           "use --impl reference for the CPU arm")
    lib = _lib.load()
    # one process per GPU: keep this rank (and the pinned staging buffers it allocates) on the
    # GPU's own NUMA node
    numa_cpus = dpc.bind_to_device_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    cfg = make_cfg(w)
    P, N, V = w["P"], w["N"], w["V"]
    Vz = V
    kern = dpc.smoothing_kernel(cfg, w["sigma"])
    taps = ops.host_taps(kern)
    params = ops.make_params(cfg, P, N, flip_y=True)

    # ---- inputs: N_INPUT_SETS seeded sets, pinned on the host, resident on the device ----
    host = [synth_inputs(w, 1000 + 17 * rank + i) for i in range(N_INPUT_SETS)]
    for h in host:
        for k in h:
            h[k] = h[k].contiguous().pin_memory()
    devin = [{k: v.to(dev) for k, v in h.items()} for h in host]
    f32 = dict(dtype=torch.float32, device=dev)
    buf = dict(tr_pc=torch.empty(P, N, 3, **f32), grid=torch.empty(P, Vz, V, V, **f32),
               bits=torch.empty(P, Vz, V, V // 32, dtype=torch.int32, device=dev),
               mask=torch.empty(P, V, V, **f32), depth=torch.empty(P, V, V, **f32),
               g_grid=torch.empty(P, Vz, V, V, **f32), g_points=torch.empty(P, N, 3, **f32),
               g_quat=torch.empty(P, 4, **f32), g_scale=torch.empty(P, **f32),
               cells=None if args.global_grid else torch.empty(
                   lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev))
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    P_ = ops._ptr
    tap_args = ops._tap_args(taps)

    def step_abi(i):
        d = devin[i % N_INPUT_SETS]
        st = lib.dpc_project_fwd(ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None,
                                 P_(d["scale"]), *tap_args, _lib.SCATTER_ATOMIC, P_(buf["tr_pc"]),
                                 P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]), P_(buf["mask"]),
                                 P_(buf["depth"]), None, None, P_(ws), ws.numel(), sptr)
        _lib.check(st, "project_fwd")
        st = lib.dpc_project_bwd(ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None,
                                 P_(d["scale"]), *tap_args, P_(buf["grid"]), P_(buf["bits"]),
                                 P_(buf["cells"]), P_(d["g_mask"]), P_(d["g_depth"]), None, None, None,
                                 P_(buf["g_grid"]), P_(buf["g_points"]), P_(buf["g_quat"]), None,
                                 None, P_(buf["g_scale"]), P_(ws), ws.numel(), sptr)
        _lib.check(st, "project_bwd")

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, join=()):
        """CUDA events on the launching stream over exactly `steps` steps; `join` = side
        streams whose work (the last step's copies) must be inside the timed region."""
        for i in range(warmup):
            fn(i)
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        for side in join:
            stream.wait_stream(side)
        e1.record(stream)
        fence()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.5)
    # ---- (1) value: device-resident inputs, C-ABI calls ----
    t_load0 = time.time()
    if args.graph:
        # the N_INPUT_SETS steps captured once into a CUDA graph and replayed
        for i in range(3):
            step_abi(i)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(stream)
        with torch.cuda.stream(cap):
            sptr_saved = sptr.value
            sptr.value = cap.cuda_stream
            with torch.cuda.graph(graph, stream=cap):
                for i in range(N_INPUT_SETS):
                    step_abi(i)
            sptr.value = sptr_saved
        stream.wait_stream(cap)
        reps = (args.steps + N_INPUT_SETS - 1) // N_INPUT_SETS

        def replay(i):
            if i % N_INPUT_SETS == 0:
                graph.replay()
        ms_abi = timed(replay, reps * N_INPUT_SETS, N_INPUT_SETS * 2) * args.steps / (reps * N_INPUT_SETS)
    else:
        ms_abi = timed(step_abi, args.steps, max(args.warmup, 3))
    # ---- (1b) informational: the same steps fed as TWO independent streams of batches (a second
    # set of buffers and a second graph on a second stream, replayed alternately): what the GPU
    # sustains when a step's kernel tails are filled by another step's kernels.  `value` stays the
    # one-stream figure.
    ms_two = None
    if args.graph and not args.global_grid:
        buf1, ws1 = buf, ws
        buf = {k: (None if v is None else torch.empty_like(v)) for k, v in buf1.items()}
        ws = torch.empty_like(ws1)
        lane_streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        for i in range(3):
            step_abi(i)
        torch.cuda.synchronize(dev)
        graph2 = torch.cuda.CUDAGraph()
        cap.wait_stream(stream)
        with torch.cuda.stream(cap):
            sptr_saved = sptr.value
            sptr.value = cap.cuda_stream
            with torch.cuda.graph(graph2, stream=cap):
                for i in range(N_INPUT_SETS):
                    step_abi(i)
            sptr.value = sptr_saved
        stream.wait_stream(cap)
        buf, ws = buf1, ws1
        lanes = (graph, graph2)

        def replay2(i):
            if i % N_INPUT_SETS == 0:
                l = (i // N_INPUT_SETS) % 2
                lane_streams[l].wait_stream(stream)
                with torch.cuda.stream(lane_streams[l]):
                    lanes[l].replay()
        reps2 = reps + (reps % 2)
        ms_two = timed(replay2, reps2 * N_INPUT_SETS, N_INPUT_SETS * 4, join=lane_streams) * args.steps / (
            reps2 * N_INPUT_SETS)
    # the timed region can be shorter than nvidia-smi's sampling period: keep
    # the same step running until the sampler has seen >= 1.5 s under load
    k = 0
    while time.time() < t_load0 + 1.5:
        step_abi(k)
        k += 1
        if k % 50 == 0:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0 + 0.2, t_load1) if sampler else None

    # ---- (2) e2e: public Python API, pinned host inputs in, results out, every step ----
    dpc.set_outputs(voxels=False, drc_probs=False)
    dpc.point_cloud._options["plane_local"] = not args.global_grid
    E2E_LANES = int(os.environ.get("DPC_E2E_LANES", "2"))

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
            # E2E_GRAPH_STEPS consecutive e2e steps (each with its own H2D and D2H copies) captured
            # once per lane with the public helpers and replayed: the eager loop is bound by the
            # host (~0.25 ms of Python / autograd-engine work per ~0.15 ms step); two lanes
            # replayed in turn keep the copy pipeline from draining at every replay boundary
            ag = dpc.AlternatingGraphs(make_step, E2E_GRAPH_STEPS, dev, lanes=E2E_LANES, warmup=2)
            # at least 8 replays (the first replay of every lane ramps its copy pipeline up), a
            # multiple of the lane count; the time is scaled to args.steps below
            reps = max(8, (args.steps + E2E_GRAPH_STEPS - 1) // E2E_GRAPH_STEPS)
            reps += reps % E2E_LANES and (E2E_LANES - reps % E2E_LANES)

            def replay_e2e(i):
                if i % E2E_GRAPH_STEPS == 0:
                    ag.replay()
            ms = timed(replay_e2e, reps * E2E_GRAPH_STEPS, 2 * E2E_LANES * E2E_GRAPH_STEPS,
                       join=ag.streams) * args.steps / (reps * E2E_GRAPH_STEPS)
            mode = "AlternatingGraphs(%d lanes x %d steps per CUDA graph)" % (E2E_LANES, E2E_GRAPH_STEPS)
            for p_ in ag.pipes:
                p_.drain()
        else:
            pipe = dpc.HostPipeline(dev, depth=3)
            ms = timed(make_step(0, pipe), args.steps, max(args.warmup, 3), join=(pipe.h2d, pipe.d2h))
            mode = "eager"
            pipe.drain()
        return ms, mode

    out_shapes = dict(mask=(P, V, V, 1), depth=(P, V, V, 1), g_points=(P, N, 3), g_quat=(P, 4),
                      g_scale=(P, 1))
    h2d = sum(host[0][k].numel() * 4 for k in ("points", "quat", "scale"))
    d2h = sum(4 * int(torch.Size(shp).numel()) for shp in out_shapes.values())
    ms_e2e, e2e_mode = measure_e2e(make_e2e(dpc.pointcloud_project_fast, host, out_shapes, h2d, d2h))

    # ---- (2b) the same step through the replica-aware API (next row f2): the host holds the
    # UN-replicated clouds (P / candidates of them), the kernels read cloud b // candidates, and
    # the cloud gradient comes back summed over the candidates -- 4x fewer point bytes each way
    R = args.candidates
    e2e_rep = None
    if R > 1 and P % R == 0:
        B = P // R
        host_rep = [{"points": h["points"][::R].contiguous().pin_memory(), "quat": h["quat"],
                     "scale": h["scale"]} for h in host]
        shapes_rep = dict(out_shapes, g_points=(B, N, 3))
        h2d_rep = sum(host_rep[0][k].numel() * 4 for k in ("points", "quat", "scale"))
        d2h_rep = sum(4 * int(torch.Size(shp).numel()) for shp in shapes_rep.values())
        ms_rep, mode_rep = measure_e2e(make_e2e(dpc.pointcloud_project_replicated, host_rep,
                                                shapes_rep, h2d_rep, d2h_rep))
        e2e_rep = {"value": world * P * args.steps / (ms_rep * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": h2d_rep, "d2h_bytes_per_step": d2h_rep,
                   "ms_per_step": ms_rep / args.steps,
                   "api": "pytorch_unsup_pc_b200.pointcloud_project_replicated: %d clouds x %d pose "
                          "candidates per step, cloud gradient summed over the candidates in the "
                          "kernels" % (B, R), "mode": mode_rep}
    # ---- (3) per-stage CUDA-event timings for the roofline ----
    stage_ms = (ctypes.c_float * len(_lib.PROFILE_STAGES))()
    d = devin[0]
    st = lib.dpc_project_profile(
        ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]), *tap_args,
        _lib.SCATTER_ATOMIC, P_(buf["tr_pc"]), P_(buf["grid"]), P_(buf["bits"]), P_(buf["cells"]),
        P_(buf["mask"]), P_(buf["depth"]), P_(d["g_mask"]), P_(d["g_depth"]), P_(buf["g_grid"]), P_(buf["g_points"]),
        P_(buf["g_quat"]), None, None, P_(buf["g_scale"]), P_(ws), ws.numel(), sptr,
        min(max(args.steps, 10), 50), stage_ms)
    _lib.check(st, "project_profile")
    fence()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    stages = dict(zip(_lib.PROFILE_STAGES, [float(x) for x in stage_ms]))
    n_chunks = lib.dpc_project_chunks(ctypes.byref(params))
    sbytes = stage_algorithmic_bytes(N, V, Vz)
    top = max((k for k in stages if k != "memset"), key=lambda k: stages[k])
    achieved = sbytes[top] * P / (stages[top] * 1e-3) / 1e9
    value = world * P * args.steps / (ms_abi * 1e-3)
    e2e_value = world * P * args.steps / (ms_e2e * 1e-3)
    step_frac = (value / world) * algorithmic_bytes(N, V, Vz) / 1e9 / peak

    # ---- (4) CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        P_cpu = 1 if args.workload == "B" else 8
        v, times = cpu_reference_rate(w, P_cpu, reps=3, warm=1, threads=threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d projections of the same workload, fwd+bwd, 1 warm-up + median of 3 "
                         "(oracle/closed_form.py: the reference's torch ops in fp64)" % P_cpu}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_abi / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": w["name"], "P_per_gpu": P, "N": N, "V": V, "Vz": Vz, "K": w["K"],
                   "sigma": w["sigma"],
                   "scatter": "global grid, atomic" if args.global_grid else "plane-local (shared memory)",
                   "l2": "no flush: %d input sets rotate and each step's grid + gradient-grid "
                         "working set (%d MiB) exceeds the 126 MB L2" % (
                             N_INPUT_SETS, 2 * P * Vz * V * V * 4 >> 20),
                   "parallelism": "projections sharded across ranks, no collective",
                   "host_binding": ("rank bound to the %d CPUs local to its GPU (NVML affinity)"
                                    % len(numa_cpus)) if numa_cpus else "none"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "api": "pytorch_unsup_pc_b200.pointcloud_project_fast + torch.autograd.grad, "
                       "host copies through pytorch_unsup_pc_b200.HostPipeline (3 streams: this "
                       "step's H2D / kernels / D2H overlap the neighbouring steps'), steps "
                       "captured and replayed with pytorch_unsup_pc_b200.AlternatingGraphs",
                "mode": e2e_mode},
        "e2e_replica_aware": e2e_rep,
        "throughput_two_streams": None if ms_two is None else {
            "value": world * P * args.steps / (ms_two * 1e-3), "unit": UNIT,
            "ms_per_step": ms_two / args.steps,
            "note": "informational: the same device-resident steps issued as two independent "
                    "streams of batches (two graphs, two streams, replayed alternately); `value` "
                    "is the one-stream figure"},
        # kernels per chunk: pose_bin (pose_cells + bin_points above 16384 points; pose_scatter on
        # the global-grid path), blur_xy, blurz_drc_fwd | drc_blurz_bwd, blur_xy, gather_pose_bwd
        # -- times the chunks the batch is split into
        # (whole job: every rank launches its own)
        "gpu_launches": (6 if args.global_grid else lib.dpc_project_kernels_per_chunk(
            ctypes.byref(params))) * n_chunks * args.steps * world,
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (None if args.global_grid else
                                 NCU_TRAFFIC_BYTES.get(args.workload, {}).get(top)),
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
                     # (timing probes)
                     "traffic_gbs": (None if (args.global_grid or NCU_TRAFFIC_BYTES.get(
                         args.workload, {}).get(top) is None) else
                         NCU_TRAFFIC_BYTES[args.workload][top] / (stages[top] * 1e-3) / 1e9),
                     "limiter": "shared-memory bandwidth / FP32 FMA issue, not HBM"},
        "roofline_step": {"algorithmic_bytes_per_projection": algorithmic_bytes(N, V, Vz),
                          "frac": step_frac, "peak": peak, "unit": "GB/s"},
        "stage_ms": stages,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="A", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--candidates", type=int, default=4,
                    help="pose candidates per cloud for the e2e_replica_aware leg (workload A: 16 x 4)")
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
