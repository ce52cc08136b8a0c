"""CPU, world_size 2 over gloo: the N>1 host logic of the path.

Each rank runs the projection of ITS samples (here with the CPU oracle -- the
test is about the sharding, the GPU kernels are covered by the -m gpu tests),
the shards are gathered and must reproduce the single-process result exactly:
the path has no cross-rank term."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form as CF
from oracle.config import default_cfg
from pytorch_unsup_pc_b200 import sharding


def test_sample_ranges_partition():
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.sample_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.sample_range(n, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    assert sharding.projection_range(16, 4, 1, 2) == (32, 64)
    with pytest.raises(ValueError):
        sharding.sample_range(4, 2, 2)
    with pytest.raises(ValueError):
        sharding.shard({"x": torch.zeros(7, 3)}, 2, 4, 0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(n_samples, replicas, N):
    g = torch.Generator().manual_seed(31)
    P = n_samples * replicas
    clouds = (torch.rand(n_samples, N, 3, generator=g) - 0.5) * 0.9
    return {"points": clouds.repeat_interleave(replicas, 0),      # tf_repeat_0 layout
            "quat": torch.randn(P, 4, generator=g),
            "scale": 0.2 + 0.8 * torch.rand(P, 1, generator=g)}


def _project(cfg, d):
    pts = d["points"].clone().requires_grad_()
    out = CF.project(cfg, pts, d["quat"], None, CF.smoothing_taps(cfg, 1.5), d["scale"])
    loss = out["proj"].sum() + 0.1 * out["proj_depth"].sum()
    (gp,) = torch.autograd.grad(loss, pts)
    return out["proj"].detach(), out["proj_depth"].detach(), gp


def _worker(rank, world, port, n_samples, replicas, N, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
        mine = sharding.shard(_inputs(n_samples, replicas, N), n_samples, replicas, rank, world)
        proj, depth, gp = _project(cfg, mine)
        # per-cloud gradient = sum over that cloud's replicas: local by construction
        g_cloud = gp.reshape(-1, replicas, N, 3).sum(1)
        parts = [None] * world
        dist.all_gather_object(parts, (proj, depth, g_cloud))
        ms = sharding.max_over_ranks(10.0 + rank)          # slowest rank wins
        if rank == 0:
            ret["proj"] = torch.cat([p[0] for p in parts])
            ret["depth"] = torch.cat([p[1] for p in parts])
            ret["g_cloud"] = torch.cat([p[2] for p in parts])
            ret["ms"] = ms
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_reproduce_single_process():
    n_samples, replicas, N, world = 3, 2, 300, 2     # uneven split: 2 + 1 samples
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_samples, replicas, N, ret), nprocs=world,
                 join=True)
        got = dict(ret)
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    proj, depth, gp = _project(cfg, _inputs(n_samples, replicas, N))
    assert torch.equal(got["proj"], proj)
    assert torch.equal(got["depth"], depth)
    assert torch.equal(got["g_cloud"], gp.reshape(n_samples, replicas, N, 3).sum(1))
    assert got["ms"] == 11.0
    assert sharding.job_rate(64, 10, 2.0, 2) == 2 * 64 * 10 / 2e-3
