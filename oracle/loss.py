"""CPU oracle of the candidate-selection projection loss.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in torch on the CPU
and in fp64 like the reference's own projections:

* ``ModelPointCloud.add_proj_loss``            models/model_pc_to.py:339-385
  (average-pool the ground-truth masks from G x G down to the prediction's
  V x V with ``nn.AvgPool2d(G // V)``, :349-356, NCHW -> NHWC :368), and
* ``ModelPointCloud.proj_loss_pose_candidates`` models/model_pc_to.py:410-440
  (replicate gt per candidate :417, per-candidate sum of squared differences
  :418-420, ``argmin`` over the candidates :421, one-hot mask :425-430,
  optional per-sample weights :431-435, ``sum(((gt-pred)*mask)**2) / BV`` :436-437).

Gradients come from torch autograd over this restatement, as the reference
gets its own.
"""
import torch


def pool_gt(masks, V):
    """masks [BV,1,G,G] -> [BV,V,V,1] (model_pc_to.py:346-356, 368)."""
    G = masks.shape[2]
    assert G >= V and G % V == 0, "GT size should not be higher than prediction size"
    if G > V:
        masks = torch.nn.functional.avg_pool2d(masks, G // V)
    return masks.permute(0, 2, 3, 1)


def proj_loss_pose_candidates(gt, pred, num_candidates, valid_samples=None):
    """gt [BV,V,V,1], pred [BV*C,V,V,1] -> (proj_loss [], min_loss [BV] int64)."""
    gt = gt.to(pred.dtype).repeat_interleave(num_candidates, 0)            # tf_repeat_0, :417
    all_loss = ((gt - pred) ** 2).sum((1, 2, 3)).reshape(-1, num_candidates)
    min_loss = all_loss.argmin(1)                                          # :421
    mask = torch.nn.functional.one_hot(min_loss, num_candidates).to(pred.dtype).reshape(-1, 1, 1, 1)
    loss_tensor = (gt - pred) * mask                                       # :430
    if valid_samples is not None:                                          # :431-435
        w = valid_samples.to(pred.dtype).repeat_interleave(num_candidates, 0).reshape(-1, 1, 1, 1)
        loss_tensor = loss_tensor * w
    return (loss_tensor ** 2).sum() / min_loss.shape[0], min_loss                  # num_samples = BV, :426


def add_proj_loss(masks, projs, num_candidates, weight_scale=1.0, valid_samples=None):
    """masks [BV,1,G,G], projs [BV*C,V,V,1] -> (total_loss [], min_loss [BV]).
    The candidate branch of add_proj_loss (:370-385): pool, select, scale."""
    gt = pool_gt(masks, projs.shape[2])
    loss, min_loss = proj_loss_pose_candidates(gt, projs, num_candidates, valid_samples)
    return loss * weight_scale, min_loss
