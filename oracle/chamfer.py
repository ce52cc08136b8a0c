"""CPU oracle of the Chamfer nearest-neighbour search.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
util/point_cloud_distance.py:25-40: ``diff = Vt - Vs`` (:33), ``dist =
sqrt(sum(diff**2, 2))`` (:34; the three squares are summed in order, verified
bit-equal to torch's reduction), ``idx = argmin(dist, 1)`` (:35, first minimum),
``proj = Vt[idx]``, ``minDist = dist[idx]`` (:38-39).

numpy, not torch: the reference evaluates on CUDA, whose ``sqrtf`` is correctly
rounded like numpy's; torch's *CPU* vectorised sqrt is 1 ulp off for ~0.7 % of
inputs (measured here on torch 2.11), which would move ``minDist`` by an ulp.
"""
import numpy as np


def point_cloud_distance(Vs, Vt, chunk=512):
    Vs = np.asarray(Vs, dtype=np.float32)
    Vt = np.asarray(Vt, dtype=np.float32)
    idx = np.empty(Vs.shape[0], dtype=np.int64)
    mind = np.empty(Vs.shape[0], dtype=np.float32)
    for s in range(0, Vs.shape[0], chunk):
        d = Vt[None, :, :] - Vs[s:s + chunk, None, :]                 # [c,M,3] fp32
        sq = d * d
        d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        dist = np.sqrt(d2)
        i = np.argmin(dist, axis=1)
        idx[s:s + chunk] = i
        mind[s:s + chunk] = dist[np.arange(dist.shape[0]), i]
    return Vt[idx], mind, idx
