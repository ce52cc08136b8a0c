#!/usr/bin/env python
"""Generate tests/golden/chamfer.npz by executing the REAL reference's
util/point_cloud_distance.py:25-40 on the CPU (needs /root/reference):
    python tests/golden/make_golden_chamfer.py
Inputs: a seeded predicted cloud and a "ground-truth" cloud that contains exact
duplicates of some points (exact distance ties: argmin must take the first)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader as RL   # noqa: E402


def make_inputs(seed=61, N=700, M=1500):
    g = torch.Generator().manual_seed(seed)
    Vs = (torch.rand(N, 3, generator=g) - 0.5)
    Vt = (torch.rand(M, 3, generator=g) - 0.5)
    Vt[M // 2:M // 2 + 200] = Vt[:200]          # duplicates later in the array
    Vs[:50] = Vt[100:150]                       # zero distances
    return Vs, Vt


def main():
    Vs, Vt = make_inputs()
    proj, dist, idx = RL.ref_point_cloud_distance(Vs, Vt)
    np.savez_compressed(os.path.join(HERE, "chamfer.npz"), Vs=Vs.numpy(), Vt=Vt.numpy(),
                        proj=proj.numpy(), minDist=dist.numpy(), idx=idx.numpy())
    print("mean dist", dist.mean().item(), "idx[:8]", idx[:8].tolist())


if __name__ == "__main__":
    main()
