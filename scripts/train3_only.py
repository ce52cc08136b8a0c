"""(torchrun) only the `workloads.train3` record of bench.py -- dev helper."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
import torch.distributed as dist

import bench
from pytorch_unsup_pc_b200 import train_step

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
args = argparse.Namespace(steps=40, warmup=5)
env = bench.Env(args, rank, world, lr)
if world > 1:
    opts = None
    if os.environ.get("DPC_NCCL_HIGH_PRIORITY"):
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=env.dev, pg_options=opts)
rec = train_step.bench(env, args)
env.fence()
if world > 1:
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps(rec))
