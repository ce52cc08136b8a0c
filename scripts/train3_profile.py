"""Dev helper: top CUDA kernels of one eager config-3 train step (torch profiler), 1 GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import train_step as TS

dev = torch.device("cuda:0")
cfg = TS.train_cfg()
kernel = dpc.smoothing_kernel(cfg, 3.0)
nets = TS.StandInNets(cfg).to(dev)
opt = torch.optim.Adam(nets.parameters(), lr=1e-4, weight_decay=1e-3, fused=True)
im, mk = TS.synth_batch(cfg, dev, 1)
for _ in range(5):
    TS.train_step(nets, opt, im, mk, cfg, kernel)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        TS.train_step(nets, opt, im, mk, cfg, kernel)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
