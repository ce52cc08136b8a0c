"""Quick per-stage timing of the projection (dev helper, not the bench)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import ops, _lib
from pytorch_unsup_pc_b200.config import default_cfg

def main(P=64, N=8000, V=64, K=21, sigma=3.0, iters=20):
    cfg = default_cfg(vox_size=V, pc_gauss_kernel_size=K)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1001)
    pts = ((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9).to(dev).requires_grad_()
    quat = torch.randn(P, 4, generator=g).to(dev).requires_grad_()
    scale = (0.2 + 0.8 * torch.rand(P, 1, generator=g)).to(dev).requires_grad_()
    kern = dpc.smoothing_kernel(cfg, sigma)
    Wp = torch.rand(P, V, V, 1, device=dev); Wd = torch.rand(P, V, V, 1, device=dev)
    dpc.set_outputs(voxels=False, drc_probs=False)
    def step():
        out = dpc.pointcloud_project_fast(cfg, pts, quat, None, None, kern, scaling_factor=scale)
        loss = (out["proj"] * Wp).sum() + 0.1 * (out["proj_depth"] * Wd).sum()
        return torch.autograd.grad(loss, [pts, quat, scale])
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"P={P} N={N} V={V} K={K} sigma={sigma}: {ms*1e3:.1f} us/step  {P/ms*1e3:.0f} proj/s (python API, fwd+bwd)")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

if __name__ == "__main__":
    a = [float(x) if "." in x else int(x) for x in sys.argv[1:]]
    main(*a)
