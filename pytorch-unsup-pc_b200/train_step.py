"""BASELINE.json configs[2]: one ``chair_unsupervised`` training step -- CNN encoder, point-cloud
decoder, pose-candidate ensemble, the differentiable projection, the candidate-selection loss and
Adam -- under ``DistributedDataParallel`` (NCCL all-reduce of the weight gradients: the path's
ONLY collective, north_star).

The networks are NOT part of this work (north_star: "The CNN encoder/decoder stays on stock
PyTorch/cuDNN"): ``StandInNets`` is a stand-in written in plain ``torch.nn`` from the layer
shapes the reference's model has (dpc/nets/img_encoder_to.py:14-64,
pc_decoder_to.py:15-56, pose_net_to.py:15-86, models/model_pc_to.py:112-125), so that the step
has the reference's compute and -- what matters for the collective -- its gradient volume.
It holds the 33.3 M parameters that receive a gradient in the shipped configuration; the
reference constructs 27.7 M more (rgb decoder heads, the pose student, the focal-length head)
that ``pc_rgb: false`` / this loss never touch -- under DDP they would need
``find_unused_parameters=True`` and would add nothing to the all-reduce but zeros.

The renderer and its loss run as the package's fused step ``project_candidates_loss``
(replica-aware: the decoder's [B,N,3] clouds are never replicated views x candidates times).
Per-rank batch = ``batch_size`` objects x ``step_size`` views (the reference's
``pool_single_view`` indexes with ``cfg.batch_size``, model_base_to.py:7-9, so the per-rank batch
is what ``cfg.batch_size`` must be under DDP).  Weak scaling.
"""
import math
import time

import torch
import torch.nn as nn

from . import losses
from .config import default_cfg
from .gauss_kernel import smoothing_kernel

TRAIN_DEFAULTS = dict(
    # dpc/resources/default_config.yaml:20-22, 38, 53, 58, 107-120, 175 + chair_unsupervised
    z_dim=1024, f_dim=16, fc_dim=1024, input_shape=[128, 128, 3], pose_candidates_num_layers=3,
    pose_predict_num_candidates=4, pc_unit_cube=True, pc_occupancy_scaling_maximum=1.0,
    step_size=4, batch_size=16, learning_rate=1e-4, weight_decay=1e-3, vox_size=64,
    pc_gauss_kernel_size=21, pc_relative_sigma=3.0, pc_num_points=8000, proj_weight=1.0,
)


def train_cfg(**overrides):
    cfg = default_cfg(**TRAIN_DEFAULTS)
    cfg.update(overrides)
    return cfg


def _init(m):
    if isinstance(m, (nn.Linear, nn.Conv2d)):
        nn.init.xavier_uniform_(m.weight)
        nn.init.constant_(m.bias, 0.01)


class StandInNets(nn.Module):
    """Encoder + decoder + scale head + pose ensemble with the reference's layer shapes.

    images [B*views,3,S,S] -> clouds [B,N,3] (first view of every object), scales [B,1],
    poses [B*views*C,4] (candidates of a view adjacent)."""

    def __init__(self, cfg):
        super().__init__()
        S, ch = int(cfg.input_shape[0]), int(cfg.input_shape[2])
        f, fc, z = int(cfg.f_dim), int(cfg.fc_dim), int(cfg.z_dim)
        act = nn.LeakyReLU()
        convs = [nn.Conv2d(ch, f, 5, stride=2, padding=2), act]
        for _ in range(int(math.log2(S / 4)) - 1):          # down to a 4 x 4 map
            convs += [nn.Conv2d(f, 2 * f, 3, stride=2, padding=1), act,
                      nn.Conv2d(2 * f, 2 * f, 3, stride=1, padding=1), act]
            f *= 2
        self.convs = nn.Sequential(*convs)
        self.fc1 = nn.Sequential(nn.Linear(f * 16, fc), act)
        self.fc2 = nn.Sequential(nn.Linear(fc, fc), act)
        self.fc3 = nn.Sequential(nn.Linear(fc, z), act)
        self.pose_fc = nn.Linear(fc, z)
        self.N = int(cfg.pc_num_points)
        self.points_fc = nn.Linear(fc, 3 * self.N)
        self.scale_fc = nn.Linear(z, 1)
        C, L = int(cfg.pose_predict_num_candidates), int(cfg.pose_candidates_num_layers)
        branches = []
        for _ in range(C):
            layers, d = [], z
            for k in range(L):
                last = k == L - 1
                layers.append(nn.Linear(d, 4 if last else 32))
                if not last:
                    layers.append(act)
                d = 32
            branches.append(nn.Sequential(*layers))
        self.pose_branches = nn.ModuleList(branches)
        self.views = int(cfg.step_size)
        self.unit_cube = bool(cfg.pc_unit_cube)
        self.scale_max = float(cfg.pc_occupancy_scaling_maximum)
        self.apply(_init)

    def forward(self, images):
        x = self.convs(images * 2 - 1)
        h2 = self.fc2(self.fc1(x.flatten(1)))
        ids = self.fc3(h2)[:: self.views]                       # the first view of every object
        pts = torch.tanh(self.points_fc(ids).reshape(-1, self.N, 3))
        if self.unit_cube:
            pts = pts / 2.0
        scale = torch.sigmoid(self.scale_fc(ids)) * self.scale_max
        pf = self.pose_fc(h2)
        poses = torch.cat([b(pf) for b in self.pose_branches], dim=1).reshape(-1, 4)
        return pts, scale, poses


def n_parameters(module):
    return sum(p.numel() for p in module.parameters())


def train_step(nets, opt, images, masks, cfg, kernel, keep_prob=1.0, seed=None,
               loss_fn=losses.project_candidates_loss):
    """One optimisation step; returns the (detached) loss.  ``nets`` may be DDP-wrapped: its
    all-reduce buckets fire as the backward reaches the parameters."""
    views, C = int(cfg.step_size), int(cfg.pose_predict_num_candidates)
    pts, scale, poses = nets(images)
    all_scale = scale.repeat_interleave(views * C, dim=0)            # [P,1], tf_repeat_0 order
    out = loss_fn(cfg, pts, poses, None, masks, kernel, scaling_factor=all_scale,
                  weight_scale=float(cfg.proj_weight), keep_prob=keep_prob, seed=seed)
    opt.zero_grad(set_to_none=True)
    out["loss"].backward()
    opt.step()
    return out["loss"].detach()


class GraphedTrainStep:
    """The whole optimisation step -- networks forward, fused projection + loss, backward, DDP's
    bucketed NCCL all-reduce, Adam -- captured ONCE into a CUDA graph and replayed.

    Eager, the step is bound by the HOST (~3.7 ms of Python / launch work for ~1.8 ms of GPU work
    at the config-3 shapes), and every DDP bucket adds host work on that critical path, so the
    collective cannot hide behind anything: 2 GPUs run at 0.85 of one.  Replayed from a graph the
    step is GPU-bound and the all-reduce buckets are graph branches that overlap the encoder's
    backward.  Follows PyTorch's recipe for DDP under whole-network capture: the DDP wrapper is
    built on a side stream, >= 11 eager iterations run before the capture, Adam is capturable."""

    def __init__(self, nets, cfg, kernel, device, world, example, warmup=11, ddp_kwargs=None,
                 comm_hook=None):
        from torch.nn.parallel import DistributedDataParallel as DDP
        self.cfg, self.kernel, self.device = cfg, kernel, device
        self.images = torch.empty_like(example[0])
        self.masks = torch.empty_like(example[1])
        self.images.copy_(example[0])
        self.masks.copy_(example[1])
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            self.model = (DDP(nets, device_ids=[device.index], **(ddp_kwargs or {}))
                          if world > 1 else nets)
            if world > 1 and comm_hook is not None:
                self.model.register_comm_hook(state=None, hook=comm_hook)
            self.opt = torch.optim.Adam(nets.parameters(), lr=float(cfg.learning_rate),
                                        weight_decay=float(cfg.weight_decay), capturable=True,
                                        fused=True)
            for _ in range(warmup):
                self.loss = train_step(self.model, self.opt, self.images, self.masks, cfg, kernel)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        # captured on the stream the DDP wrapper was built and warmed up on: its AccumulateGrad
        # nodes (DDP keeps references to them) belong to that stream
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = train_step(self.model, self.opt, self.images, self.masks, cfg, kernel)

    def __call__(self, images, masks):
        """Copies the batch into the captured step's input buffers and replays it; returns the
        loss tensor of the captured step (overwritten by the next call)."""
        self.images.copy_(images, non_blocking=True)
        self.masks.copy_(masks, non_blocking=True)
        self.graph.replay()
        return self.loss


def synth_batch(cfg, device, seed):
    """Synthetic per-rank batch: images rand [B*views,3,S,S], Bernoulli(0.5) masks [B*views,1,S,S]
    (SURVEY.md 8d)."""
    g = torch.Generator().manual_seed(seed)
    BV, S = int(cfg.batch_size) * int(cfg.step_size), int(cfg.input_shape[0])
    images = torch.rand(BV, int(cfg.input_shape[2]), S, S, generator=g)
    masks = (torch.rand(BV, 1, S, S, generator=g) > 0.5).float()
    return images.to(device), masks.to(device)


def bench(env, args):
    """bench.py's `workloads.train3` record: the DDP train step at this world size, replayed from
    a CUDA graph (``GraphedTrainStep``; `eager` holds the same step issued from Python), the
    share of the step spent in the projection + loss, and the cost of the gradient all-reduce:
    stand-alone, and exposed (step with it minus the same graphed step without it)."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    dev, world, rank = env.dev, env.world, env.rank
    cfg = train_cfg()
    kernel = smoothing_kernel(cfg, float(cfg.pc_relative_sigma))
    torch.backends.cudnn.benchmark = True         # stock cuDNN, its own algorithm search
    batches = [synth_batch(cfg, dev, 4000 + 31 * rank + i) for i in range(3)]
    steps = max(10, min(args.steps, 40))
    P = int(cfg.batch_size) * int(cfg.step_size) * int(cfg.pose_predict_num_candidates)
    ddp_kw = dict(gradient_as_bucket_view=True)

    def fresh_nets():
        torch.manual_seed(1234)                   # same initial weights on every rank
        return StandInNets(cfg).to(dev)

    # (1) the step replayed from a CUDA graph, DDP inside the capture
    nets = fresh_nets()
    n_par = n_parameters(nets)
    graphed = GraphedTrainStep(nets, cfg, kernel, dev, world, batches[0], ddp_kwargs=ddp_kw)
    ms = env.timed(lambda i: graphed(*batches[i % 3]), steps, 5)
    loss = float(graphed.loss)
    rec = {"value": world * P * steps / (ms * 1e-3), "unit": "projections/s",
           "steps_per_s": steps / (ms * 1e-3),
           "samples_per_s": world * int(cfg.batch_size) * steps / (ms * 1e-3),
           "ms_per_step": ms / steps, "steps": steps, "loss": loss, "scaling": "weak",
           "mode": "whole step (networks, projection + loss, backward, DDP all-reduce buckets, Adam) "
                   "captured once into a CUDA graph and replayed",
           "config": {"workload": "chair_unsupervised train step per GPU: batch %d x %d views x %d "
                                  "candidates, %d points -> %d^3, K=%d sigma=%.1f, stand-in nets in "
                                  "torch.nn (%.1f M parameters), Adam, fp32" % (
                                      cfg.batch_size, cfg.step_size, cfg.pose_predict_num_candidates,
                                      cfg.pc_num_points, cfg.vox_size, cfg.pc_gauss_kernel_size,
                                      cfg.pc_relative_sigma, n_par / 1e6),
                      "parallelism": "torch.nn.parallel.DistributedDataParallel over NCCL, %d rank(s), "
                                     "gradient_as_bucket_view, 25 MB buckets" % world,
                      "parameters": n_par, "allreduce_bytes_per_step": 4 * n_par if world > 1 else 0}}
    del graphed, nets
    # (2) the same step issued eagerly from Python (host-bound)
    nets = fresh_nets()
    model = DDP(nets, device_ids=[dev.index], **ddp_kw) if world > 1 else nets
    opt = torch.optim.Adam(nets.parameters(), lr=float(cfg.learning_rate),
                           weight_decay=float(cfg.weight_decay), fused=True)
    ms_e = env.timed(lambda i: train_step(model, opt, *batches[i % 3], cfg, kernel), steps, 5)
    rec["eager"] = {"ms_per_step": ms_e / steps, "value": world * P * steps / (ms_e * 1e-3),
                    "note": "the same DDP step launched from Python every iteration: bound by the "
                            "host, so every bucket's launch work is on the critical path"}
    del model, opt, nets
    # (3) the projection + loss alone on the same shapes (fused step, eager, same API call)
    pts = ((torch.rand(int(cfg.batch_size), int(cfg.pc_num_points), 3, device=dev) - 0.5) * 0.9).requires_grad_()
    quat = torch.randn(P, 4, device=dev).requires_grad_()
    scl = (0.2 + 0.8 * torch.rand(P, 1, device=dev)).requires_grad_()

    def proj_only(i):
        out = losses.project_candidates_loss(cfg, pts, quat, None, batches[i % 3][1], kernel,
                                             scaling_factor=scl)
        torch.autograd.grad(out["loss"], [pts, quat, scl])
    ms_p = env.timed(proj_only, steps, 5)
    rec["projection_ms_per_step"] = ms_p / steps
    rec["projection_share"] = ms_p / ms
    if world > 1:
        # (4) the collective: the graphed step without it (every rank trains alone), and alone
        solo = GraphedTrainStep(fresh_nets(), cfg, kernel, dev, 1, batches[0])
        ms_ns = env.timed(lambda i: solo(*batches[i % 3]), steps, 5)
        del solo
        # the same captured step with DDP's stock bf16 compression hook: the buckets travel as
        # bf16 (half the bytes) and are decompressed into the fp32 gradients -- a different
        # gradient-averaging arithmetic (the reference has no distributed training to compare
        # with), reported as a variant, never as `value`
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        half = GraphedTrainStep(fresh_nets(), cfg, kernel, dev, world, batches[0], ddp_kwargs=ddp_kw,
                                comm_hook=default_hooks.bf16_compress_hook)
        ms_bf = env.timed(lambda i: half(*batches[i % 3]), steps, 5)
        del half
        rec["sync_variants"] = {
            "fp32_buckets (value)": {"ms_per_step": ms / steps, "efficiency_vs_no_collective": ms_ns / ms},
            "bf16_compress_hook": {"ms_per_step": ms_bf / steps, "efficiency_vs_no_collective": ms_ns / ms_bf,
                                   "note": "torch.distributed.algorithms.ddp_comm_hooks.default_hooks."
                                           "bf16_compress_hook: 66.5 MB instead of 133 MB per step"}}
        flat = torch.empty(n_par, dtype=torch.float32, device=dev)
        ms_ar = env.timed(lambda i: dist.all_reduce(flat), steps, 3)
        rec["allreduce"] = {
            "collective": "NCCL all-reduce over NVLink / NVSwitch of the %.1f MB of fp32 weight "
                          "gradients, issued by DistributedDataParallel per 25 MB bucket as the backward "
                          "produces them (graph branches of the captured step)" % (4 * n_par / 1e6),
            "standalone_ms": ms_ar / steps,
            "standalone_busbw_gbs": 4 * n_par * 2 * (world - 1) / world / (ms_ar / steps * 1e-3) / 1e9,
            "step_without_allreduce_ms": ms_ns / steps,
            "exposed_ms": max(0.0, (ms - ms_ns) / steps),
            "overlapped_fraction": max(0.0, min(1.0, 1.0 - (ms - ms_ns) / max(ms_ar, 1e-9)))}
    del batches
    # (graphs that hold captured NCCL work are gone by now: dist.destroy_process_group() does not
    # return while one is alive)
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return rec
