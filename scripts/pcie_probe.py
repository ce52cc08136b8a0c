"""Host<->device copy rates of this box at the e2e step's sizes (dev helper): H2D alone, D2H alone,
both directions at once -- the ceiling of bench.py's `e2e` leg."""
import torch

dev = torch.device("cuda:0")
h_in = torch.empty(64 * 8000 * 3, dtype=torch.float32).pin_memory()
h_out = torch.empty(2 * 64 * 64 * 64 + 64 * 8000 * 3, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty_like(h_in, device=dev), torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h, n=300):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    for _ in range(n):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1)
    cur.wait_stream(s2)
    e1.record(cur)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, a, b in (("H2D alone", 1, 0), ("D2H alone", 0, 1), ("both", 1, 1)):
    run(a, b, 20)
    us = run(a, b)
    gb = lambda t, on: ("%.1f GB/s" % (t.numel() * 4 / us / 1e3)) if on else "-"
    print("%-10s %.1f us/step   H2D %s   D2H %s" % (name, us, gb(h_in, a), gb(h_out, b)))
