"""CPU-side cost of one e2e step (dev helper): wall time per step of the
Python API loop with and without host copies, plus a cProfile of the loop."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import pytorch_unsup_pc_b200 as dpc  # noqa: E402

w = bench.WORKLOADS["A"]
cfg = bench.make_cfg(w)
dev = torch.device("cuda:0")
kern = dpc.smoothing_kernel(cfg, w["sigma"])
host = bench.synth_inputs(w, 1000)
for k in host:
    host[k] = host[k].contiguous().pin_memory()
d = {k: v.to(dev) for k, v in host.items()}
dpc.set_outputs(voxels=False, drc_probs=False)
P, N, V = w["P"], w["N"], w["V"]
out_host = dict(mask=torch.empty(P, V, V, 1).pin_memory(), depth=torch.empty(P, V, V, 1).pin_memory(),
                g_points=torch.empty(P, N, 3).pin_memory(), g_quat=torch.empty(P, 4).pin_memory(),
                g_scale=torch.empty(P, 1).pin_memory())
pipe = dpc.HostPipeline(dev, depth=3)


def step_dev():
    pts = d["points"].detach().requires_grad_()
    quat = d["quat"].detach().requires_grad_()
    scale = d["scale"].detach().requires_grad_()
    out = dpc.pointcloud_project_fast(cfg, pts, quat, None, None, kern, scaling_factor=scale)
    return torch.autograd.grad([out["proj"], out["proj_depth"]], [pts, quat, scale],
                               [d["g_mask"], d["g_depth"]]), out


def step_e2e():
    din = pipe.upload({"points": host["points"], "quat": host["quat"], "scale": host["scale"]})
    pts = din["points"].detach().requires_grad_()
    quat = din["quat"].detach().requires_grad_()
    scale = din["scale"].detach().requires_grad_()
    out = dpc.pointcloud_project_fast(cfg, pts, quat, None, None, kern, scaling_factor=scale)
    gp, gq, gs = torch.autograd.grad([out["proj"], out["proj_depth"]], [pts, quat, scale],
                                     [d["g_mask"], d["g_depth"]])
    pipe.download({"mask": out["proj"], "depth": out["proj_depth"], "g_points": gp, "g_quat": gq,
                   "g_scale": gs}, out_host)


def copies_only():
    din = pipe.upload({"points": host["points"], "quat": host["quat"], "scale": host["scale"]})
    pipe.download({"g_points": din["points"]}, out_host)


for name, fn in (("device-only", step_dev), ("e2e", step_e2e), ("copies-only", copies_only)):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-12s issue %.1f us/step   complete %.1f us/step" % (name, (t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))

pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step_e2e()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
