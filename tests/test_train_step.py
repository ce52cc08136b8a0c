"""configs[2] host logic on the CPU: the stand-in networks' shapes and parameter count, and the
DistributedDataParallel wiring of ``train_step`` at world size 2 over gloo (a stand-in loss in
place of the CUDA projection: the test is about the collective -- after a step on DIFFERENT
per-rank batches every rank must hold the same weights, equal to a single-process step on the
mean gradient)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorch_unsup_pc_b200 import train_step as TS


def test_stand_in_nets_shapes_and_parameter_count():
    cfg = TS.train_cfg()
    nets = TS.StandInNets(cfg)
    # encoder 8.52 M (convs 1.18 M + fc 7.34 M incl. pose_fc), decoder 1024 -> 3 x 8000 24.6 M,
    # 4 pose branches 0.14 M, scale head 1 k  (the reference's unused rgb / student /
    # focal heads are not built)
    n = TS.n_parameters(nets)
    assert 33.0e6 < n < 33.6e6, n
    small = TS.train_cfg(input_shape=[32, 32, 3], pc_num_points=50, batch_size=2, step_size=2)
    nets = TS.StandInNets(small)
    pts, scale, poses = nets(torch.rand(4, 3, 32, 32))
    assert pts.shape == (2, 50, 3) and scale.shape == (2, 1) and poses.shape == (4 * 4, 4)
    assert float(pts.abs().max()) <= 0.5 and 0 < float(scale.min()) and float(scale.max()) < 1


def _stand_in_loss(cfg, pts, poses, trans, masks, kernel, scaling_factor=None, weight_scale=1.0,
                   keep_prob=1.0, seed=None):
    return {"loss": (pts ** 2).mean() + (poses ** 2).mean() + scaling_factor.mean() * masks.mean()}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        cfg = TS.train_cfg(input_shape=[32, 32, 3], pc_num_points=50, batch_size=2, step_size=2)
        torch.manual_seed(7)
        nets = TS.StandInNets(cfg)
        model = torch.nn.parallel.DistributedDataParallel(nets)
        opt = torch.optim.SGD(nets.parameters(), lr=0.1)
        images, masks = TS.synth_batch(cfg, "cpu", 100 + rank)
        TS.train_step(model, opt, images, masks, cfg, None, loss_fn=_stand_in_loss)
        ret[rank] = torch.cat([p.detach().reshape(-1) for p in nets.parameters()]).numpy()
    finally:
        dist.destroy_process_group()


def test_ddp_train_step_world2_gloo():
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        flats = [torch.from_numpy(ret[r]) for r in range(2)]
    assert torch.equal(flats[0], flats[1])                       # the ranks agree after the step
    # single process, mean of the two ranks' gradients: the same weights
    cfg = TS.train_cfg(input_shape=[32, 32, 3], pc_num_points=50, batch_size=2, step_size=2)
    torch.manual_seed(7)
    nets = TS.StandInNets(cfg)
    opt = torch.optim.SGD(nets.parameters(), lr=0.1)
    opt.zero_grad()
    for rank in range(2):
        images, masks = TS.synth_batch(cfg, "cpu", 100 + rank)
        pts, scale, poses = nets(images)
        loss = _stand_in_loss(cfg, pts, poses, None, masks, None,
                              scaling_factor=scale.repeat_interleave(8, 0))["loss"]
        (loss / 2).backward()
    opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in nets.parameters()])
    assert torch.allclose(flats[0], ref, rtol=1e-5, atol=1e-7)
