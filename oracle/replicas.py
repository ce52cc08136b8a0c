"""Oracle of next-row f2: cloud replication + point dropout around the projection.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

Restates, with plain torch/numpy on the CPU,
  * ``tf_repeat_0``           models/model_pc_to.py:47-56
  * ``pc_point_dropout``      util/point_cloud_to.py:269-295 (sampler :275-283, gather
                              ``select_3d`` :266-267)
  * the call order of ``ModelPointCloud.forward`` / ``compute_projection``
    (models/model_pc_to.py:302-306, 254-265): replicate by views, replicate by candidates,
    drop points from every copy, project.
Gradients: torch autograd through ``repeat`` and the index gather (sum over the replicas,
zero for dropped points) -- the reference's own mechanism.

Pinned against the reference executed live (tests/test_replicas.py::
test_oracle_matches_live_reference) and against tests/golden/replicas.npz.
"""
import numpy as np
import torch

from . import closed_form as CF


def tf_repeat_0(x, num):
    """[B, ...] -> [B*num, ...], every row repeated `num` times in place
    (model_pc_to.py:47-56: unsqueeze(1), repeat, reshape)."""
    shape = list(x.shape)
    return x.unsqueeze(1).repeat([1, num] + [1] * (len(shape) - 1)).reshape([-1] + shape[1:])


def sample_indices(batch, num_points, keep_prob):
    """The reference sampler (point_cloud_to.py:275-283): np.random.choice without
    replacement per row, consuming numpy's GLOBAL RNG in row order; [batch, M, 2] int64 of
    (row, point index) pairs."""
    M = int(num_points * keep_prob)
    rows = []
    for k in range(batch):
        ind = np.random.choice(num_points, M, replace=False)
        rows.append(np.stack([np.full_like(ind, k), ind], axis=1)[None])
    return torch.from_numpy(np.concatenate(rows, 0).astype(np.int64))


def select_3d(data, indices):
    """data[P,N,C] at indices[P,M,2] -> [P,M,C] (point_cloud_to.py:266-267)."""
    return data[indices[:, :, 0], indices[:, :, 1]]


def pc_point_dropout(points, rgb, indices):
    """point_cloud_to.py:284-294 with the sampled indices given."""
    out = select_3d(points, indices)
    out_rgb = None if rgb is None else select_3d(rgb, indices)
    return out, out_rgb


def project_replicated(cfg, point_cloud, quat, replicas, indices=None, translation=None,
                       kernel=None, scale=None, focal=None):
    """point_cloud [B,N,3], quat [B*replicas,4] -> the projection dict of
    ``closed_form.project`` on the replicated, dropped-out clouds."""
    pts = tf_repeat_0(point_cloud, replicas)
    if indices is not None:
        pts, _ = pc_point_dropout(pts, None, indices)
    return CF.project(cfg, pts, quat, translation, kernel, scale, focal)
