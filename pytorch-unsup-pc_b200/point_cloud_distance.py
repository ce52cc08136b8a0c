"""Drop-in for the reference's util/point_cloud_distance.py (SURVEY.md 8f row
f4): the nearest-neighbour search behind the Chamfer evaluation
(run/eval_chamfer_to.py:24-44), as one brute-force CUDA kernel instead of
[VsN,VtN,3] temporaries (csrc/chamfer.cu)."""
import torch

from . import _lib, ops


def point_cloud_distance(Vs, Vt):
    """For each point in Vs [VsN,3] the closest point in Vt [VtN,3]
    (point_cloud_distance.py:25-40): returns (proj [VsN,3], minDist [VsN],
    idx [VsN] int64).  fp32 (fp64 inputs are converted); no gradient -- the
    reference only calls it under evaluation."""
    src = ops._f32(Vs, "Vs")
    tgt = ops._f32(Vt, "Vt")
    if src.dim() != 2 or src.shape[1] != 3 or tgt.dim() != 2 or tgt.shape[1] != 3:
        raise ValueError("Vs and Vt must be [N,3] and [M,3], got %s and %s"
                         % (tuple(src.shape), tuple(tgt.shape)))
    if tgt.device != src.device:
        raise ValueError("Vs and Vt must be on the same device")
    src, tgt = src.detach(), tgt.detach()
    N, M = src.shape[0], tgt.shape[0]
    dev = src.device
    proj = torch.empty(N, 3, dtype=torch.float32, device=dev)
    min_dist = torch.empty(N, dtype=torch.float32, device=dev)
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    ws = ops._scratch("nn", 8 * N, dev)
    with ops._on_device(dev):
        st = _lib.load().dpc_point_cloud_distance(N, M, ops._ptr(src), ops._ptr(tgt), ops._ptr(proj),
                                                  ops._ptr(min_dist), ops._ptr(idx), ops._ptr(ws),
                                                  ws.numel(), ops._stream(dev))
    _lib.check(st, "point_cloud_distance")
    return proj, min_dist, idx


def chamfer_distances(pred, gt):
    """(mean pred->gt distance, mean gt->pred distance): what
    run/eval_chamfer_to.py:112-125 accumulates per view."""
    _, d_pg, _ = point_cloud_distance(pred, gt)
    _, d_gp, _ = point_cloud_distance(gt, pred)
    return d_pg.double().mean(), d_gp.double().mean()
