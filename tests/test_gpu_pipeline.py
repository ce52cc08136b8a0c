"""GPU: HostPipeline (three-stream host<->device overlap) returns exactly what
the single-stream path returns."""
import pytest
import torch

from oracle.config import default_cfg

pytestmark = pytest.mark.gpu


def test_host_pipeline_matches_single_stream():
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    kern = dpc.smoothing_kernel(cfg, 1.5)
    P, N, V = 4, 2000, 32
    g = torch.Generator().manual_seed(3)
    sets = [{"points": ((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9).pin_memory(),
             "quat": torch.randn(P, 4, generator=g).pin_memory()} for _ in range(5)]
    w = torch.rand(P, V, V, 1, generator=g).to(dev)

    def step(d):
        pts = d["points"].detach().requires_grad_()
        q = d["quat"].detach().requires_grad_()
        with dpc.options(deterministic=True):          # bit-exact, so torch.equal below is fair
            out = dpc.pointcloud_project_fast(cfg, pts, q, None, None, kern)
        gp, gq = torch.autograd.grad((out["proj"] * w).sum(), [pts, q])
        return {"mask": out["proj"], "g_points": gp, "g_quat": gq}

    ref = []
    for s in sets:
        r = step({k: v.to(dev) for k, v in s.items()})
        ref.append({k: v.detach().cpu() for k, v in r.items()})
    pipe = dpc.HostPipeline(dev, depth=2)
    host = [{"mask": torch.empty(P, V, V, 1).pin_memory(), "g_points": torch.empty(P, N, 3).pin_memory(),
             "g_quat": torch.empty(P, 4).pin_memory()} for _ in sets]
    for s, h in zip(sets, host):
        pipe.download(step(pipe.upload(s)), h)
    pipe.drain()
    for r, h in zip(ref, host):
        for k in r:
            assert torch.equal(r[k], h[k]), k
    assert pipe.h2d_bytes == (P * N * 3 + P * 4) * 4
    with pytest.raises(ValueError):
        pipe.upload({"points": torch.zeros(2, 3)})     # not pinned


def test_graphed_steps_replay_matches_eager():
    """A captured graph of e2e steps (copies + forward + backward) replays to the same
    host results as the eager loop, and picks up NEW data written into the staging buffers."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    kern = dpc.smoothing_kernel(cfg, 1.5)
    P, N, V, S = 4, 2000, 32, 4
    g = torch.Generator().manual_seed(9)
    stage = [{"points": ((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9).pin_memory(),
              "quat": torch.randn(P, 4, generator=g).pin_memory()} for _ in range(S)]
    host = [{"mask": torch.zeros(P, V, V, 1).pin_memory(), "g_points": torch.zeros(P, N, 3).pin_memory()}
            for _ in range(S)]
    w = torch.rand(P, V, V, 1, generator=g).to(dev)
    pipe = dpc.HostPipeline(dev, depth=2)

    def step(k):
        d = pipe.upload(stage[k])
        pts = d["points"].detach().requires_grad_()
        q = d["quat"].detach().requires_grad_()
        with dpc.options(deterministic=True, voxels=False, drc_probs=False):
            out = dpc.pointcloud_project_fast(cfg, pts, q, None, None, kern)
        (gp,) = torch.autograd.grad((out["proj"] * w).sum(), [pts])
        pipe.download({"mask": out["proj"], "g_points": gp}, host[k])

    def eager():
        for k in range(S):
            step(k)
        pipe.drain()
        return [{n: t.clone() for n, t in h.items()} for h in host]

    ref = eager()
    gs = dpc.GraphedSteps(step, S, dev, pipe=pipe)
    for h in host:
        for t in h.values():
            t.zero_()
    gs.replay()
    torch.cuda.synchronize()
    for r, h in zip(ref, host):
        for n in r:
            assert torch.equal(r[n], h[n]), n
    # new batch written into the SAME staging buffers
    for s_ in stage:
        s_["points"].copy_((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9)
    gs.replay()
    torch.cuda.synchronize()
    got = [{n: t.clone() for n, t in h.items()} for h in host]
    ref2 = eager()
    for r, h in zip(ref2, got):
        for n in r:
            assert torch.equal(r[n], h[n]), n
    assert not torch.equal(ref[0]["mask"], ref2[0]["mask"])


def test_alternating_graphs_match_eager():
    """Two lanes of captured steps replayed in turn on two streams: every lane's host results equal
    the eager loop's, replay after replay."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    kern = dpc.smoothing_kernel(cfg, 1.5)
    P, N, V, S, L = 4, 1500, 32, 3, 2
    g = torch.Generator().manual_seed(19)
    stage = [[{"points": ((torch.rand(P, N, 3, generator=g) - 0.5) * 0.9).pin_memory(),
               "quat": torch.randn(P, 4, generator=g).pin_memory()} for _ in range(S)] for _ in range(L)]
    host = [[{"mask": torch.zeros(P, V, V, 1).pin_memory(), "g_points": torch.zeros(P, N, 3).pin_memory()}
             for _ in range(S)] for _ in range(L)]
    w = torch.rand(P, V, V, 1, generator=g).to(dev)

    def make_step(lane, pipe):
        def step(k):
            d = pipe.upload(stage[lane][k])
            pts = d["points"].detach().requires_grad_()
            q = d["quat"].detach().requires_grad_()
            with dpc.options(voxels=False, drc_probs=False):     # the default path is reproducible
                out = dpc.pointcloud_project_fast(cfg, pts, q, None, None, kern)
            (gp,) = torch.autograd.grad((out["proj"] * w).sum(), [pts])
            pipe.download({"mask": out["proj"], "g_points": gp}, host[lane][k])
        return step

    def eager():
        res = []
        for lane in range(L):
            pipe = dpc.HostPipeline(dev)
            step = make_step(lane, pipe)
            for k in range(S):
                step(k)
            pipe.drain()
            res.append([{n: t.clone() for n, t in h.items()} for h in host[lane]])
        return res

    ref = eager()
    ag = dpc.AlternatingGraphs(make_step, S, dev, lanes=L)
    for rep in range(2):
        for lane in range(L):
            for h in host[lane]:
                for t in h.values():
                    t.zero_()
        for _ in range(L):
            ag.replay()
        ag.join()
        torch.cuda.synchronize()
        for lane in range(L):
            for r, h in zip(ref[lane], host[lane]):
                for n in r:
                    assert torch.equal(r[n], h[n]), (rep, lane, n)
