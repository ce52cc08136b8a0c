"""Dev helper (torchrun, N ranks): step time of the config-3 train step under several ways of
all-reducing the weight gradients -- DDP with different bucket sizes / static_graph, one flat
all-reduce after the backward, no collective at all."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import train_step as TS

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
cfg = TS.train_cfg()
kernel = dpc.smoothing_kernel(cfg, 3.0)
batches = [TS.synth_batch(cfg, dev, 4000 + 31 * rank + i) for i in range(3)]


def timed(fn, steps=40, warm=8):
    for i in range(warm):
        fn(i)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def variant(name, **ddp_kw):
    torch.manual_seed(1)
    nets = TS.StandInNets(cfg).to(dev)
    opt = torch.optim.Adam(nets.parameters(), lr=1e-4, weight_decay=1e-3, fused=True)
    if name == "none":
        model = nets
    elif name == "flat":
        model = nets
        params = [p for p in nets.parameters()]
        flat = torch.zeros(sum(p.numel() for p in params), device=dev)
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
    else:
        model = DDP(nets, device_ids=[lr], **ddp_kw)

    def step(i):
        im, mk = batches[i % 3]
        views, C = cfg.step_size, cfg.pose_predict_num_candidates
        pts, scale, poses = model(im)
        out = dpc.project_candidates_loss(cfg, pts, poses, None, mk, kernel,
                                          scaling_factor=scale.repeat_interleave(views * C, 0))
        if name == "flat":
            flat.zero_()
            out["loss"].backward()
            dist.all_reduce(flat)
            flat.div_(world)
        else:
            opt.zero_grad(set_to_none=True)
            out["loss"].backward()
        opt.step()
    ms = timed(step)
    if rank == 0:
        print("%-28s %.3f ms/step" % (name + str(ddp_kw or ""), ms), flush=True)


def graphed(**ddp_kw):
    torch.manual_seed(1)
    nets = TS.StandInNets(cfg).to(dev)
    g = TS.GraphedTrainStep(nets, cfg, kernel, dev, world if not ddp_kw.pop("single", False) else 1,
                            batches[0], ddp_kwargs=ddp_kw)

    def step(i):
        g(*batches[i % 3])
    ms = timed(step)
    if rank == 0:
        print("%-28s %.3f ms/step  loss %.4f" % ("graphed" + str(ddp_kw or ""), ms, float(g.loss)), flush=True)


torch.backends.cudnn.benchmark = True
graphed(single=True)
graphed(gradient_as_bucket_view=True)
graphed(gradient_as_bucket_view=True, bucket_cap_mb=100)
dist.destroy_process_group()
