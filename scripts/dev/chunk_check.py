import sys, os
sys.path.insert(0, "/root/repo")
import torch
import pytorch_unsup_pc_b200 as dpc
from oracle.config import default_cfg
cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
P, N = 4, 500
pts = ((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9).to(dev).requires_grad_()
q = torch.randn(P, 4, generator=g).to(dev).requires_grad_()
s = (0.2 + 0.8 * torch.rand(P, 1, generator=g)).to(dev).requires_grad_()
kern = dpc.smoothing_kernel(cfg, 1.5)
out = dpc.pointcloud_project_fast(cfg, pts, q, None, None, kern, scaling_factor=s)
loss = out["proj"].sum() + out["proj_depth"].sum()
gr = torch.autograd.grad(loss, [pts, q, s])
torch.cuda.synchronize()
print("ok", float(loss), [float(x.abs().sum()) for x in gr])
