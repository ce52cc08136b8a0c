"""Row f4: nearest neighbour of the Chamfer evaluation."""
import os

import numpy as np
import pytest
import torch

from golden.make_golden_chamfer import make_inputs
from oracle import chamfer as OC

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chamfer.npz")))


def _ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def test_oracle_matches_reference_golden():
    """idx and proj exactly; minDist to 1 ulp (torch's CPU sqrt is not correctly rounded,
    the reference's CUDA sqrtf and numpy's are -- oracle/chamfer.py)."""
    Vs, Vt = make_inputs()
    assert np.array_equal(Vs.numpy(), GOLD["Vs"]) and np.array_equal(Vt.numpy(), GOLD["Vt"])
    proj, dist, idx = OC.point_cloud_distance(GOLD["Vs"], GOLD["Vt"])
    assert np.array_equal(idx, GOLD["idx"])
    assert np.array_equal(proj, GOLD["proj"])
    assert _ulp_diff(dist, GOLD["minDist"]).max() <= 1
    assert (idx[:50] == np.arange(100, 150)).all() and (dist[:50] == 0).all()   # first of the duplicates


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(),
                    reason="reference tree not mounted")
@pytest.mark.parametrize("i", range(6))
def test_oracle_matches_live_reference_sweep(i):
    """The numpy oracle against the reference's own point_cloud_distance executed live
    (util/point_cloud_distance.py:25-40) over ragged sizes, clustered clouds, exact duplicates
    and lattice points (many exactly tied distances: the FIRST minimum must win): idx and proj
    equal, minDist within 1 ulp (torch's vectorised CPU sqrt against the correctly rounded one)."""
    from oracle import ref_loader as RL
    g = torch.Generator().manual_seed(6100 + i)
    N, M = [(1, 1), (7, 300), (257, 129), (600, 1000), (64, 64), (300, 513)][i]
    Vs = torch.rand(N, 3, generator=g) - 0.5
    Vt = torch.rand(M, 3, generator=g) - 0.5
    if i == 3:                                   # duplicates among the targets and exact hits
        Vt[500:600] = Vt[100:200]
        Vs[:40] = Vt[100:140]
    if i == 4:                                   # a lattice: every source has several equidistant targets
        Vs = torch.round(Vs * 4) / 4
        Vt = torch.round(Vt * 4) / 4
    if i == 5:
        Vs, Vt = Vs * 0.05, Vt * 0.05 + 0.3     # far, tight clusters: near-ties after rounding
    r_proj, r_dist, r_idx = RL.ref_point_cloud_distance(Vs, Vt)
    proj, dist, idx = OC.point_cloud_distance(Vs.numpy(), Vt.numpy())
    assert np.array_equal(idx, r_idx.numpy())
    assert np.array_equal(proj, r_proj.numpy())
    assert _ulp_diff(dist, r_dist.numpy().astype(np.float32)).max() <= 1


@pytest.mark.gpu
def test_cuda_is_bit_exact_against_the_oracle():
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    proj, dist, idx = dpc.point_cloud_distance(torch.from_numpy(GOLD["Vs"]).to(dev),
                                               torch.from_numpy(GOLD["Vt"]).to(dev))
    o_proj, o_dist, o_idx = OC.point_cloud_distance(GOLD["Vs"], GOLD["Vt"])
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(dist.cpu().numpy(), o_dist)
    assert np.array_equal(proj.cpu().numpy(), o_proj)
    assert np.array_equal(idx.cpu().numpy(), GOLD["idx"])


@pytest.mark.gpu
@pytest.mark.parametrize("N,M", [(1, 1), (5, 3000), (8000, 10000), (1030, 1025)])
def test_cuda_sizes_and_target_slices(N, M):
    """Ragged sizes, more than one target slice, a near-degenerate cloud (many near-ties)."""
    import pytorch_unsup_pc_b200 as dpc
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + M)
    Vs = torch.rand(N, 3, generator=g)
    Vt = torch.rand(M, 3, generator=g)
    Vt[:, 2] = (Vt[:, 2] * 4).round() / 4              # coarse z: many equal / nearly equal distances
    proj, dist, idx = dpc.point_cloud_distance(Vs.to(dev), Vt.to(dev))
    o_proj, o_dist, o_idx = OC.point_cloud_distance(Vs.numpy(), Vt.numpy())
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(dist.cpu().numpy(), o_dist)
    a, b = dpc.chamfer_distances(Vs.to(dev), Vt.to(dev))
    assert abs(a.item() - o_dist.astype(np.float64).mean()) < 1e-12


def test_cpu_tensors_are_refused():
    import pytorch_unsup_pc_b200 as dpc
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpc.point_cloud_distance(torch.zeros(4, 3), torch.zeros(5, 3))
