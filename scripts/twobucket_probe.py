"""Dev probe (torchrun, 2+ ranks): hand-scheduled two-bucket gradient all-reduce for the config-3
train step, eager and captured into a CUDA graph.  Prints progress so a hang can be located."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pytorch_unsup_pc_b200 as dpc
from pytorch_unsup_pc_b200 import train_step as TS

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True


def say(*a):
    if rank == 0:
        print(*a, file=sys.stderr, flush=True)


cfg = TS.train_cfg()
kernel = dpc.smoothing_kernel(cfg, 3.0)
batches = [TS.synth_batch(cfg, dev, 4000 + 31 * rank + i) for i in range(3)]
torch.manual_seed(1)
nets = TS.StandInNets(cfg).to(dev)
params = list(nets.parameters())
head_ids = {id(p) for m in (nets.points_fc, nets.scale_fc, nets.pose_branches) for p in m.parameters()}
order = [p for p in params if id(p) not in head_ids] + [p for p in params if id(p) in head_ids]
flat = torch.zeros(sum(p.numel() for p in params), device=dev)
off = 0
for p in order:
    p.grad = flat[off:off + p.numel()].view_as(p)
    off += p.numel()
n_trunk = sum(p.numel() for p in order if id(p) not in head_ids)
trunk, head = flat[:n_trunk], flat[n_trunk:]
state = {"pending": 0, "work": []}
MODE = os.environ.get("TB_MODE", "hook")


def hook(_p):
    state["pending"] -= 1
    if state["pending"] == 0 and MODE == "hook":
        state["work"].append(dist.all_reduce(head, async_op=True))


for p in order:
    if id(p) in head_ids:
        p.register_post_accumulate_grad_hook(hook)
opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-3, capturable=True, fused=True)
images, masks = batches[0][0].clone(), batches[0][1].clone()
views, C = cfg.step_size, cfg.pose_predict_num_candidates
out_loss = [None]


def step():
    state["pending"], state["work"] = len(head_ids), []
    flat.zero_()
    pts, scale, poses = nets(images)
    out = dpc.project_candidates_loss(cfg, pts, poses, None, masks, kernel,
                                      scaling_factor=scale.repeat_interleave(views * C, 0))
    (out["loss"] / world).backward()
    if MODE != "hook":
        state["work"].append(dist.all_reduce(head, async_op=True))
    state["work"].append(dist.all_reduce(trunk, async_op=True))
    for w in state["work"]:
        w.wait()
    opt.step()
    out_loss[0] = out["loss"].detach()


def timed(fn, steps=40, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(5):
        step()
        say("eager step", i, "issued")
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
say("eager warm-up done, loss", float(out_loss[0]))
say("eager two-bucket (%s): %.3f ms/step" % (MODE, timed(step)))
g = torch.cuda.CUDAGraph()
say("capturing")
with torch.cuda.graph(g, stream=side):
    step()
say("captured")


def replay():
    g.replay()


say("graphed two-bucket (%s): %.3f ms/step, loss %.4f" % (MODE, timed(replay), float(out_loss[0])))
# a CUDA graph that holds captured NCCL work must be gone before the process group is torn down
# (destroy_process_group() otherwise never returns)
del g
torch.cuda.synchronize()
dist.destroy_process_group()
