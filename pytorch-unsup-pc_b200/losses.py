"""Candidate-selection projection loss: drop-in for
``ModelPointCloud.add_proj_loss`` / ``proj_loss_pose_candidates``
(models/model_pc_to.py:339-385, 410-440), fused into one forward and one
backward kernel (csrc/candidate_loss.cu): AvgPool2d of the ground truth,
per-candidate squared error, argmin, one-hot-masked loss and its gradient.
"""
import ctypes

import torch

from . import _lib, ops
from . import point_cloud as _pc


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


def project_candidates_loss(cfg, point_cloud, transform, predicted_translation, masks, kernel=None,
                            scaling_factor=None, focal_length=None, weight_scale=1.0,
                            valid_samples=None, keep_prob=1.0, indices=None, seed=None):
    """The renderer and its loss as ONE differentiable step -- what ``ModelPointCloud.forward``
    + ``get_loss`` run between the decoder and ``loss.backward()`` (model_pc_to.py:302-331:
    tf_repeat_0 of the clouds over views x candidates, pc_point_dropout,
    pointcloud_project_fast; :339-385, 410-440: AvgPool2d of the masks, per-candidate error,
    argmin, one-hot-masked loss).

    ``point_cloud`` [B,N,3] the decoder's clouds (un-replicated), ``transform`` [P,4] the
    P = B x views x candidates poses in tf_repeat_0 order, ``masks`` [BV,1,G,G] the ground
    truth of the BV = B x views views, ``scaling_factor`` / ``focal_length`` /
    ``predicted_translation`` per projection ([P,1] / [P,1] / [P,3]) as the model replicates
    them; ``keep_prob`` / ``indices`` / ``seed`` as in ``pointcloud_project_replicated``.

    Returns {"loss" [] (differentiable), "min_loss" [BV] int64, "all_loss" [BV,C],
    "projs" [P,V,V,1] (detached), "dropout_indices"}.  Equal to
    ``add_proj_loss(cfg, {"masks": masks}, {"projs": pointcloud_project_replicated(...)["proj"]},
    weight_scale)`` to rounding; the gradient of every losing candidate is exactly zero
    (:425-430), so the backward kernels run over the winning candidates only, and dL/dmask is
    formed inside the ray kernel instead of travelling through memory."""
    _pc._check_quaternion_cfg(cfg)
    if getattr(cfg, "ptn_max_projection", False):
        raise NotImplementedError("ptn_max_projection is broken in the reference "
                                  "(point_cloud_to.py:234,242) and not supported")
    if getattr(cfg, "pc_gauss_filter_gt", False):
        raise NotImplementedError("pc_gauss_filter_gt is 'Not implemented' in the reference "
                                  "(model_pc_to.py:357-358)")
    C = int(cfg.pose_predict_num_candidates)
    if C <= 1:
        raise NotImplementedError("single-candidate branch: the reference evaluates "
                                  "nn.MSELoss(gt - pred), which fails (model_pc_to.py:376)")
    pts = ops._f32(point_cloud, "point_cloud")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError("point_cloud must be [B,N,3], got %s" % (tuple(pts.shape),))
    B, N_src, _ = pts.shape
    quat = ops._f32(transform, "transform")
    if quat.dim() != 2 or quat.shape[1] != 4 or quat.shape[0] % (B * C) != 0:
        raise ValueError("transform must be [B*views*%d,4] for %d clouds, got %s"
                         % (C, B, tuple(quat.shape)))
    P = quat.shape[0]
    BV = P // C
    gt = ops._f32(masks, "masks")
    G = gt.shape[-2] if gt.dim() == 4 and gt.shape[-1] == 1 else gt.shape[-1]
    gt = gt.reshape(-1, G, G)
    V = int(cfg.vox_size)
    if gt.shape[0] != BV:
        raise ValueError("masks: %d images, expected B*views=%d" % (gt.shape[0], BV))
    if G < V or G % V:
        raise ValueError("GT size should not be higher than prediction size")   # :347
    weights = None
    if valid_samples is not None and getattr(cfg, "variable_num_views", False):
        weights = ops._f32(valid_samples, "valid_samples").reshape(-1)
        if weights.numel() != BV:
            raise ValueError("valid_samples: expected %d values" % BV)
    sel = None
    if indices is not None:
        sel = _pc._selection(indices, P, N_src, pts.device)
    elif keep_prob != 1:
        M = int(N_src * keep_prob)
        if not 1 <= M <= N_src:
            raise ValueError("keep_prob=%r keeps %d of %d points" % (keep_prob, M, N_src))
        sel = ops.dropout_indices(P, N_src, M, _pc._new_seed() if seed is None else seed, pts.device)
    N = N_src if sel is None else sel.shape[1]
    trans = ops._f32(predicted_translation, "predicted_translation", (P, 3))
    scale = _pc._vec(scaling_factor, "scaling_factor", P)
    focal = _pc._vec(focal_length, "focal_length", P)
    params = ops.make_params(cfg, P, N, flip_y=True)
    taps = ops.host_taps(kernel)
    loss, min_idx, all_loss, mask = ops.RenderLossFn.apply(
        pts, quat, trans, focal, scale, gt, weights, params, taps, C, float(weight_scale),
        _pc._scatter_mode(), _pc._options["plane_local"], P // B, N_src, sel)
    return {"loss": loss, "min_loss": min_idx, "all_loss": all_loss, "projs": mask.unsqueeze(-1),
            "dropout_indices": sel}
