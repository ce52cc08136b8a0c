"""Go/no-go probe for a tensor-core blur pass: an optimistic bound.

One 21-tap blur pass over a batch of 64^3 grids is out[rows, 64] = in[rows, 64 + 20] x T[84, 64]
with T banded Toeplitz (rows = P * 64 * 64).  fp32 accuracy from TF32 tensor cores takes the
3 x TF32 split (hi*hi + hi*lo + lo*hi): three GEMMs per pass.  This times the LIBRARY TF32 GEMM
(cuBLAS through torch.matmul, the dense band included -- a hand-written tcgen05 kernel could skip
~2/3 of the band's zero blocks but would also have to split the operands and stage the planes)
and the same product in fp32 SIMT, next to the FFMA2 kernels' measured time per pass.
    python scripts/tf32_blur_probe.py"""
import torch

dev = torch.device("cuda:0")
P, V, K = 64, 64, 21
rows = P * V * V
a = torch.rand(rows, V + K - 1, device=dev)
taps = torch.softmax(-((torch.arange(K) - 10.0) ** 2) / 18.0, 0)
T = torch.zeros(V + K - 1, V)
for j in range(V):
    T[j:j + K, j] = taps
T = T.to(dev)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


ref = (a.double() @ T.double())
torch.backends.cuda.matmul.allow_tf32 = False
t_fp32 = timed(lambda: a @ T)
torch.backends.cuda.matmul.allow_tf32 = True
t_tf32 = timed(lambda: a @ T)
e1 = float(((a @ T).double() - ref).abs().max())
# 3 x TF32: split both operands into a TF32-exact high part and the remainder
def split(x):
    hi = (x.view(torch.int32) & -8192).view(torch.float32)       # keep 10 mantissa bits
    return hi, x - hi
ah, al = split(a)
th, tl = split(T)
t_3x = timed(lambda: ah @ th + (ah @ tl + al @ th))
e3 = float(((ah @ th + (ah @ tl + al @ th)).double() - ref).abs().max())
torch.backends.cuda.matmul.allow_tf32 = False
e0 = float(((a @ T).double() - ref).abs().max())
print("one blur pass over %d x 64^3 (rows=%d, K=%d): fp32 SIMT GEMM %.1f us (err %.1e) | TF32 GEMM %.1f us "
      "(err %.1e) | 3xTF32 (three GEMMs + adds) %.1f us (err %.1e)" % (P, rows, V + K - 1, t_fp32, e0, t_tf32, e1, t_3x, e3))
