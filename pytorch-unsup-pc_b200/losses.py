"""Candidate-selection projection loss: drop-in for
``ModelPointCloud.add_proj_loss`` / ``proj_loss_pose_candidates``
(models/model_pc_to.py:339-385, 410-440), fused into one forward and one
backward kernel (csrc/candidate_loss.cu): AvgPool2d of the ground truth,
per-candidate squared error, argmin, one-hot-masked loss and its gradient.
"""
import ctypes

import torch

from . import _lib, ops


class CandidateLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, masks, projs, weights, num_candidates, weight_scale):
        lib = _lib.load()
        dev = projs.device
        BV, G = masks.shape[0], masks.shape[-1]
        V = projs.shape[1]
        C = int(num_candidates)
        f32 = dict(dtype=torch.float32, device=dev)
        all_loss = torch.empty(BV, C, **f32)
        view_loss = torch.empty(BV, **f32)
        min_idx = torch.empty(BV, dtype=torch.int64, device=dev)
        with ops._on_device(dev):
            st = lib.dpc_candidate_loss_fwd(BV, C, V, G, ops._ptr(masks), ops._ptr(projs),
                                            ops._ptr(weights), ops._ptr(all_loss), ops._ptr(min_idx),
                                            ops._ptr(view_loss), ops._stream(dev))
        _lib.check(st, "candidate_loss_fwd")
        ctx.save_for_backward(masks, projs, weights, min_idx)
        ctx.dims = (BV, C, V, G)
        ctx.coeff = float(weight_scale) / BV
        ctx.mark_non_differentiable(min_idx, all_loss)
        return view_loss.sum() * ctx.coeff, min_idx, all_loss

    @staticmethod
    def backward(ctx, g_total, _g_idx, _g_all):
        lib = _lib.load()
        masks, projs, weights, min_idx = ctx.saved_tensors
        BV, C, V, G = ctx.dims
        dev = projs.device
        g_total = ops._f32(g_total, "grad_output").reshape(())
        g_pred = torch.empty_like(projs)
        with ops._on_device(dev):
            st = lib.dpc_candidate_loss_bwd(BV, C, V, G, ops._ptr(masks), ops._ptr(projs),
                                            ops._ptr(weights), ops._ptr(min_idx), ops._ptr(g_total),
                                            ctypes.c_float(ctx.coeff), ops._ptr(g_pred),
                                            ops._stream(dev))
        _lib.check(st, "candidate_loss_bwd")
        return None, g_pred, None, None, None


def _candidate_loss(masks, projs, num_candidates, weight_scale, valid_samples):
    projs = ops._f32(projs, "projs")
    if projs.dim() == 4 and projs.shape[-1] == 1:
        projs = projs.squeeze(-1)
    if projs.dim() != 3 or projs.shape[1] != projs.shape[2]:
        raise ValueError("projs must be [BV*C,V,V,1], got %s" % (tuple(projs.shape),))
    C = int(num_candidates)
    if C < 1 or projs.shape[0] % C:
        raise ValueError("projs: %d images is not a multiple of num_candidates=%d" % (projs.shape[0], C))
    BV, V = projs.shape[0] // C, projs.shape[1]
    masks = ops._f32(masks, "masks")
    G = masks.shape[-2] if masks.dim() == 4 and masks.shape[-1] == 1 else masks.shape[-1]
    masks = masks.reshape(-1, G, G)
    if masks.shape[0] != BV:
        raise ValueError("masks: %d images, expected BV=%d" % (masks.shape[0], BV))
    if G < V or G % V:
        raise ValueError("GT size should not be higher than prediction size")   # :347
    weights = None
    if valid_samples is not None:
        weights = ops._f32(valid_samples, "valid_samples").reshape(-1)
        if weights.numel() != BV:
            raise ValueError("valid_samples: expected %d values" % BV)
    total, min_idx, all_loss = CandidateLossFn.apply(masks, projs, weights, C, float(weight_scale))
    return total, min_idx, all_loss


def proj_loss_pose_candidates(cfg, gt, pred, inputs=None):
    """gt [BV,V,V,1], pred [BV*C,V,V,1] -> (proj_loss [], min_loss [BV])
    (model_pc_to.py:410-440).  ``inputs['valid_samples']`` is used when
    ``cfg.variable_num_views`` is set (:431-435)."""
    vs = inputs["valid_samples"] if (getattr(cfg, "variable_num_views", False) and inputs) else None
    total, min_idx, _ = _candidate_loss(gt, pred, cfg.pose_predict_num_candidates, 1.0, vs)
    return total, min_idx


def add_proj_loss(cfg, inputs, outputs, weight_scale):
    """``inputs['masks']`` [BV,1,G,G], ``outputs['projs']`` [BV*C,V,V,1] ->
    (total_loss [], min_loss [BV]) (model_pc_to.py:339-385): the AvgPool2d of
    the masks is fused into the kernel.  Unlike the reference, ``inputs`` is
    not modified (:369 stores the pooled masks back)."""
    if getattr(cfg, "pc_gauss_filter_gt", False):
        raise NotImplementedError("pc_gauss_filter_gt is 'Not implemented' in the reference "
                                  "(model_pc_to.py:357-358)")
    if getattr(cfg, "pose_predictor_student", False):
        raise NotImplementedError("the student loss (model_pc_to.py:442-) is outside this path; "
                                  "call proj_loss_pose_candidates and add it to the result")
    C = int(cfg.pose_predict_num_candidates)
    if C <= 1:
        raise NotImplementedError("single-candidate branch: the reference evaluates "
                                  "nn.MSELoss(gt - pred), which fails (model_pc_to.py:376)")
    vs = inputs.get("valid_samples") if getattr(cfg, "variable_num_views", False) else None
    total, min_idx, _ = _candidate_loss(inputs["masks"], outputs["projs"], C, weight_scale, vs)
    return total, min_idx
