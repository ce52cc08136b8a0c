"""Per-stage CUDA-event timings of one projection step through dpc_project_profile (dev helper
for kernel A/B runs: DPC_B200_LIB=<variant .so> python scripts/stage_time.py [A|B] [iters] [box]).
`box` < 0.9 shrinks the synthetic clouds to a cube of that side (an object that fills part of
the frustum: planes of depth no point touches).  DPC_STAGE_SORTED=1 runs the deterministic
(sort-then-segment) mode; the printed digest covers every output and gradient bit for bit."""
import ctypes
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import _lib, ops


def main(workload="A", iters=30, box=0.9, P=0):
    w = dict(bench.WORKLOADS[workload])
    if P:
        w["P"] = int(P)
    cfg = bench.make_cfg(w)
    lib = _lib.load()
    dev = torch.device("cuda:0")
    P, N, V = w["P"], w["N"], w["V"]
    Vz = V
    taps = ops.host_taps(dpc.smoothing_kernel(cfg, w["sigma"]))
    params = ops.make_params(cfg, P, N, flip_y=True)
    d = {k: v.to(dev) for k, v in bench.synth_inputs(w, 1000).items()}
    d["points"] = d["points"] * (box / 0.9)
    f32 = dict(dtype=torch.float32, device=dev)
    buf = dict(tr_pc=torch.empty(P, N, 3, **f32), grid=torch.empty(P, Vz, V, V, **f32),
               bits=torch.empty(P, Vz, V, V // 32, dtype=torch.int32, device=dev),
               mask=torch.empty(P, V, V, **f32), depth=torch.empty(P, V, V, **f32),
               g_grid=torch.empty(P, Vz, V, V, **f32), g_points=torch.empty(P, N, 3, **f32),
               g_quat=torch.empty(P, 4, **f32), g_scale=torch.empty(P, **f32),
               cells=torch.empty(lib.dpc_cells_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev))
    ws = torch.empty(lib.dpc_workspace_bytes(ctypes.byref(params)), dtype=torch.uint8, device=dev)
    sptr = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P_ = ops._ptr
    stage_ms = (ctypes.c_float * len(_lib.PROFILE_STAGES))()
    mode = _lib.SCATTER_SORTED if os.environ.get("DPC_STAGE_SORTED") else _lib.SCATTER_ATOMIC
    for it in (3, iters):
        st = lib.dpc_project_profile(
            ctypes.byref(params), P_(d["points"]), P_(d["quat"]), None, None, P_(d["scale"]),
            *ops._tap_args(taps), mode, P_(buf["tr_pc"]), P_(buf["grid"]), P_(buf["bits"]),
            P_(buf["cells"]), P_(buf["mask"]), P_(buf["depth"]), P_(d["g_mask"]), P_(d["g_depth"]),
            P_(buf["g_grid"]), P_(buf["g_points"]), P_(buf["g_quat"]), None, None, P_(buf["g_scale"]),
            P_(ws), ws.numel(), sptr, it, stage_ms)
        _lib.check(st, "profile")
    torch.cuda.synchronize()
    t = {k: round(float(v) * 1e3, 1) for k, v in zip(_lib.PROFILE_STAGES, stage_ms)}
    chk = [float(buf[k].double().abs().sum()) for k in ("mask", "depth", "g_points", "g_quat", "g_scale")]
    h = hashlib.sha256()
    for k in ("tr_pc", "mask", "depth", "g_points", "g_quat", "g_scale"):
        h.update(buf[k].cpu().numpy().tobytes())
    print(os.path.basename(_lib.LIB_PATH), workload, "sorted" if mode else "atomic", "P", P, "box %.2f" % box, t,
          "sum %.1f" % sum(t.values()), "chk", " ".join("%.6e" % c for c in chk), "sha", h.hexdigest()[:16])


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["A"]), *[int(x) for x in sys.argv[2:3]], *[float(x) for x in sys.argv[3:4]],
         *[int(x) for x in sys.argv[4:5]])
