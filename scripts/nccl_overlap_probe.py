"""Dev helper (torchrun, >= 2 ranks): does an NCCL all-reduce captured in a CUDA graph overlap an
independent chain of compute kernels captured in the same graph?"""
import os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
flat = torch.zeros(33_000_000, device=dev)
a = torch.randn(2048, 2048, device=dev)
b = torch.randn(2048, 2048, device=dev)


def compute(n=8):
    x = a
    for _ in range(n):
        x = x @ b
        x = x / x.abs().max()
    return x


def capture(fn):
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    return g


def timed(g, n=200):
    for _ in range(5):
        g.replay()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def both_async():
    w = dist.all_reduce(flat, async_op=True)
    compute()
    w.wait()


def both_serial():
    dist.all_reduce(flat)
    compute()


r = {}
r["compute"] = timed(capture(compute))
r["allreduce"] = timed(capture(lambda: dist.all_reduce(flat)))
r["serial"] = timed(capture(both_serial))
r["async"] = timed(capture(both_async))


class Eager:
    def __init__(self, fn):
        self.fn = fn

    def replay(self):
        self.fn()


r["eager_compute"] = timed(Eager(compute))
r["eager_serial"] = timed(Eager(both_serial))
r["eager_async"] = timed(Eager(both_async))
if rank == 0:
    print({k: round(v, 3) for k, v in r.items()}, flush=True)
dist.destroy_process_group()
