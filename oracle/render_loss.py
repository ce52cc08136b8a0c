"""Oracle of the fused renderer + candidate-selection loss step.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

The composition ``ModelPointCloud.forward`` + ``get_loss`` run between the decoder and
``loss.backward()`` (models/model_pc_to.py:302-331 and :339-385, 410-440), out of the three
pinned restatements: ``oracle.replicas`` (tf_repeat_0 + pc_point_dropout), ``oracle.closed_form``
(the projection) and ``oracle.loss`` (AvgPool2d + candidate selection).  Pinned against the
reference's own composition by tests/golden/render_loss.npz (made by
tests/golden/make_golden_render_loss.py from the live reference).
"""
from . import loss as OL
from . import replicas as OR


def project_candidates_loss(cfg, point_cloud, quat, masks, num_candidates, kernel=None, scale=None,
                            translation=None, focal=None, weight_scale=1.0, valid_samples=None,
                            indices=None):
    """point_cloud [B,N,3], quat [P,4], masks [BV,1,G,G] -> (loss, min_loss [BV], proj [P,V,V,1]).
    ``indices``: the dropout selection as [P,M,2] (row, point) pairs, or None."""
    replicas = quat.shape[0] // point_cloud.shape[0]
    out = OR.project_replicated(cfg, point_cloud, quat, replicas, indices, translation, kernel,
                                scale, focal)
    loss, min_loss = OL.add_proj_loss(masks, out["proj"], num_candidates, weight_scale, valid_samples)
    return loss, min_loss, out["proj"]
