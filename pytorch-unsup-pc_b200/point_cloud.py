"""Drop-in mirror of the reference's util/point_cloud_to.py (hot-path functions).

Same names, argument order and output layouts as the reference; outputs are
fp32 (the reference's are fp64 because of a float64 constant in
quaternion_conjugate; its callers only subtract and sum them).
"""
import contextlib

import torch

from . import _lib, ops

# plane_local: build/gather every Z-plane in shared memory (default); False keeps
# the raw grid in global memory (memset + atomic scatter, grid gather) -- same
# results to rounding, kept for A/B measurements
# validate_indices: range- and duplicate-check a user-supplied dropout selection (one small
# reduction; one host sync when the selection lives on the device)
_options = {"voxels": True, "drc_probs": True, "deterministic": False, "plane_local": True,
            "validate_indices": True}


def set_outputs(voxels=None, drc_probs=None):
    """Choose whether ``pointcloud_project_fast`` materialises the two large,
    rarely-consumed outputs (``voxels`` [P,Vz,V,V,1], ``drc_probs``
    [Vz+1,P,V,V,1]).  Default: both on, exactly like the reference.  The
    reference model only consumes proj / proj_depth / drc_probs
    (model_pc_to.py:266-269) and drc_probs only when ``drc_weight > 0``."""
    if voxels is not None:
        _options["voxels"] = bool(voxels)
    if drc_probs is not None:
        _options["drc_probs"] = bool(drc_probs)


def set_deterministic(flag=True):
    """Use the sort-then-segment scatter: bit-exact run to run (the default
    scatter uses fp32 reductions whose order varies)."""
    _options["deterministic"] = bool(flag)


@contextlib.contextmanager
def options(**kw):
    old = dict(_options)
    try:
        for k, v in kw.items():
            if k not in _options:
                raise KeyError(k)
            _options[k] = bool(v)
        yield
    finally:
        _options.update(old)


def _scatter_mode():
    return _lib.SCATTER_SORTED if _options["deterministic"] else _lib.SCATTER_ATOMIC


def _check_quaternion_cfg(cfg):
    if not getattr(cfg, "pose_quaternion", True):
        raise NotImplementedError(
            "pose_quaternion: false (4x4 camera matrices) is not supported; the reference's "
            "matrix branch itself fails (point_cloud_to.py:153, UnboundLocalError)")


def _vec(t, name, P):
    """[P,1] / [P] -> contiguous [P] (scaling_factor, focal_length)."""
    if t is None:
        return None
    t = ops._f32(t, name)
    if t.numel() != P:
        raise ValueError("%s: expected %d values, got shape %s" % (name, P, tuple(t.shape)))
    return t.reshape(P)


def pc_perspective_transform(cfg, point_cloud, transform, predicted_translation=None,
                             focal_length=None):
    """point_cloud [P,N,3], transform [P,4] quaternion -> [P,N,3] in (z,y,x)
    order (point_cloud_to.py:118-178)."""
    _check_quaternion_cfg(cfg)
    pts = ops._f32(point_cloud, "point_cloud")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError("point_cloud must be [P,N,3], got %s" % (tuple(pts.shape),))
    P, N, _ = pts.shape
    quat = ops._f32(transform, "transform", (P, 4))
    trans = ops._f32(predicted_translation, "predicted_translation", (P, 3))
    focal = _vec(focal_length, "focal_length", P)
    params = ops.make_params(cfg, P, N)
    return ops.PoseFn.apply(pts, quat, trans, focal, params)


def pointcloud2voxels3d_fast(cfg, pc, rgb):
    """pc [P,N,3] (already transformed) -> (voxels [P,Vz,V,V], None)
    (point_cloud_to.py:10-87).  Points outside [-0.5,0.5]^3 are dropped."""
    pts = ops._f32(pc, "pc")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError("pc must be [P,N,3], got %s" % (tuple(pts.shape),))
    params = ops.make_params(cfg, pts.shape[0], pts.shape[1])
    voxels = ops.ScatterFn.apply(pts, params, _scatter_mode())
    if rgb is None:
        return voxels, None
    # the rgb part of the TF original (point_cloud.py:111-121; the torch port's :61-69 is broken):
    # voxels_rgb [P,Vz,V,V,C], the same trilinear weights times the point's features
    col = _features(rgb, pts)
    tr = pts.detach() if getattr(cfg, "pc_rgb_stop_points_gradient", False) else pts
    planar = ops.FeatGridFn.apply(tr, col, params, None, False)
    return voxels, planar.permute(0, 2, 3, 4, 1)


def _features(rgb, pts):
    col = ops._f32(rgb, "rgb")
    if col.dim() != 3 or col.shape[:2] != pts.shape[:2] or not 1 <= col.shape[2] <= 4:
        raise ValueError("rgb must be [P,N,C] with 1 <= C <= 4 for points %s, got %s"
                         % (tuple(pts.shape), tuple(col.shape)))
    return col


def convolve_rgb(cfg, voxels_rgb, kernel):
    """voxels_rgb [P,Vz,V,V,C] -> same: the three 1-D blurs on every channel separately
    (point_cloud_to.py:106-114; TF point_cloud.py:148-154)."""
    vox = ops._f32(voxels_rgb, "voxels_rgb")
    if vox.dim() != 5:
        raise ValueError("voxels_rgb must be [P,Vz,V,V,C], got %s" % (tuple(vox.shape),))
    P, Vz, V, V2, C = vox.shape
    params = ops.make_params(cfg, P * C, 0)
    if (params.Vz, params.V, params.V) != (Vz, V, V2):
        raise ValueError("voxels_rgb shape %s does not match cfg grid %s"
                         % (tuple(vox.shape), (params.Vz, params.V, params.V)))
    planar = vox.permute(0, 4, 1, 2, 3).reshape(P * C, Vz, V, V).contiguous()
    out = ops.BlurFn.apply(planar, params, ops.host_taps(kernel))
    return out.reshape(P, C, Vz, V, V).permute(0, 2, 3, 4, 1)


def smoothen_voxels3d(cfg, voxels, kernel):
    """voxels [P,1,Vz,V,V], kernel = [k1,k2,k3] -> same shape
    (point_cloud_to.py:90-103)."""
    vox = ops._f32(voxels, "voxels")
    if vox.dim() != 5 or vox.shape[1] != 1:
        raise ValueError("voxels must be [P,1,Vz,V,V], got %s" % (tuple(vox.shape),))
    P, _, Vz, V, V2 = vox.shape
    params = ops.make_params(cfg, P, 0)
    if (params.Vz, params.V, params.V) != (Vz, V, V2):
        raise ValueError("voxels shape %s does not match cfg grid %s"
                         % (tuple(vox.shape), (params.Vz, params.V, params.V)))
    out = ops.BlurFn.apply(vox.reshape(P, Vz, V, V), params, ops.host_taps(kernel))
    return out.reshape(P, 1, Vz, V, V)


def _new_seed():
    # one draw from torch's default CPU generator: follows torch.manual_seed, no device sync
    return int(torch.empty((), dtype=torch.int64).random_().item())


def _selection(indices, P, N_src, device):
    """Validate a user-supplied dropout selection: [P,M] or the reference sampler's
    [P,M,2] (batch index, point index) layout (point_cloud_to.py:275-283)."""
    sel = torch.as_tensor(indices)
    if sel.dim() == 3 and sel.shape[-1] == 2:
        sel = sel[..., 1]
    if sel.dim() != 2 or sel.shape[0] != P or not 1 <= sel.shape[1] <= N_src:
        raise ValueError("indices must be [P,M] (or the reference's [P,M,2]) with P=%d and "
                         "1 <= M <= %d, got %s" % (P, N_src, tuple(sel.shape)))
    if sel.is_floating_point() or sel.dtype == torch.bool:
        raise TypeError("indices must be an integer tensor")
    if _options["validate_indices"]:
        # an index outside the cloud would be an out-of-bounds access in the kernels, and a
        # repeated one would keep a single slot of the backward's inverse map (the gradient of
        # the other would be lost): refuse both.  The reference's sampler draws without
        # replacement (point_cloud_to.py:279), so it never produces duplicates.
        srt = torch.sort(sel.long(), dim=1).values
        bad = torch.stack([srt[:, 0].min() < 0, srt[:, -1].max() >= N_src,
                           (srt[:, 1:] == srt[:, :-1]).any()]).tolist()
        if bad[0] or bad[1]:
            raise ValueError("indices must lie in [0, %d)" % N_src)
        if bad[2]:
            raise ValueError("indices must be distinct within every row (sampling without "
                             "replacement, point_cloud_to.py:279)")
    return sel.to(device=device, dtype=torch.int32).contiguous()


def pc_point_dropout(points, rgb, keep_prob, seed=None, indices=None):
    """points [P,N,3] (+ rgb [P,N,C]) -> ([P,M,3], [P,M,C] or None), M = int(N * keep_prob):
    an independent random M-subset of every cloud (point_cloud_to.py:269-295), sampled and
    gathered on the device -- the reference samples with numpy on the host and gathers by
    advanced indexing.  ``indices`` ([P,M] or the reference's [P,M,2]) replaces the sampler;
    ``seed`` makes it reproducible (default: a draw from torch's generator)."""
    pts = ops._f32(points, "points")
    if pts.dim() != 3:
        raise ValueError("points must be [P,N,C], got %s" % (tuple(pts.shape),))
    P, N, _ = pts.shape
    if indices is None:
        M = int(N * keep_prob)
        if not 1 <= M <= N:
            raise ValueError("keep_prob=%r keeps %d of %d points" % (keep_prob, M, N))
        sel = ops.dropout_indices(P, N, M, _new_seed() if seed is None else seed, pts.device)
    else:
        sel = _selection(indices, P, N, pts.device)
    out = ops.SelectPointsFn.apply(pts, sel, 1)
    out_rgb = None
    if rgb is not None:
        col = ops._f32(rgb, "rgb")
        if col.dim() != 3 or col.shape[:2] != pts.shape[:2] or col.shape[2] > 4:
            raise ValueError("rgb must be [P,N,C<=4] like points, got %s" % (tuple(col.shape),))
        out_rgb = ops.SelectPointsFn.apply(col, sel, 1)
    return out, out_rgb


def pointcloud_project_replicated(cfg, point_cloud, transform, predicted_translation, all_rgb,
                                  kernel=None, scaling_factor=None, focal_length=None,
                                  keep_prob=1.0, indices=None, seed=None):
    """``pointcloud_project_fast`` on the UN-replicated clouds (next row f2).

    The reference materialises every predicted cloud ``step_size x num_candidates`` times
    (tf_repeat_0, model_pc_to.py:47-56, 302-306), drops points from every copy with a host-side
    sampler (pc_point_dropout, :254-258) and projects the copies.  Here ``point_cloud`` is the
    decoder's [B,N,3] tensor, ``transform`` holds the P = B x replicas poses in tf_repeat_0
    order (the replicas of a cloud adjacent) and the kernels read cloud ``b // replicas``
    through the dropout selection: neither copy exists, and the gradient comes back already
    summed over the replicas, [B,N,3].  Outputs are those of
    ``pointcloud_project_fast(cfg, pc_point_dropout(tf_repeat_0(point_cloud, replicas)), ...)``
    plus ``dropout_indices`` ([P,M] int32, or None when nothing is dropped).

    ``keep_prob`` < 1 samples M = int(N * keep_prob) points per projection on the device;
    ``indices`` ([P,M], distinct per row) supplies the selection instead."""
    _check_quaternion_cfg(cfg)
    if getattr(cfg, "ptn_max_projection", False):
        raise NotImplementedError("ptn_max_projection is broken in the reference "
                                  "(point_cloud_to.py:234,242) and not supported")
    pts = ops._f32(point_cloud, "point_cloud")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError("point_cloud must be [B,N,3], got %s" % (tuple(pts.shape),))
    B, N_src, _ = pts.shape
    quat = ops._f32(transform, "transform")
    if quat.dim() != 2 or quat.shape[1] != 4 or quat.shape[0] % B != 0:
        raise ValueError("transform must be [B*replicas,4] for %d clouds, got %s"
                         % (B, tuple(quat.shape)))
    P = quat.shape[0]
    sel = None
    if indices is not None:
        sel = _selection(indices, P, N_src, pts.device)
    elif keep_prob != 1:
        M = int(N_src * keep_prob)
        if not 1 <= M <= N_src:
            raise ValueError("keep_prob=%r keeps %d of %d points" % (keep_prob, M, N_src))
        sel = ops.dropout_indices(P, N_src, M, _new_seed() if seed is None else seed, pts.device)
    N = N_src if sel is None else sel.shape[1]
    trans = ops._f32(predicted_translation, "predicted_translation", (P, 3))
    scale = _vec(scaling_factor, "scaling_factor", P)
    focal = _vec(focal_length, "focal_length", P)
    col = None if all_rgb is None else _features(all_rgb, pts)      # [B,N,C], un-replicated too
    params = ops.make_params(cfg, P, N, flip_y=True)
    taps = ops.host_taps(kernel)
    mask, depth, tr_pc, voxels, probs = ops.project(
        pts, quat, trans, focal, scale, params, taps,
        _options["voxels"], _options["drc_probs"] or col is not None, _scatter_mode(),
        _options["plane_local"], (P // B, N_src, sel))
    voxels_rgb = proj_rgb = None
    if col is not None:
        # the features of every projection's surviving points (they are C floats per point: this
        # copy is the one replica tensor that does get materialised)
        if sel is not None:
            col = ops.SelectPointsFn.apply(col, sel, P // B)
        else:
            col = col.repeat_interleave(P // B, dim=0)
        voxels_rgb, proj_rgb = _project_features(cfg, tr_pc, col.contiguous(), probs, taps, P, N)
    return {
        "proj": mask.unsqueeze(-1),
        "voxels": None if voxels is None else voxels.unsqueeze(-1),
        "tr_pc": tr_pc,
        "voxels_rgb": voxels_rgb,
        "proj_rgb": proj_rgb,
        "drc_probs": None if (probs is None or not _options["drc_probs"]) else probs.unsqueeze(-1),
        "proj_depth": depth.unsqueeze(-1),
        "dropout_indices": sel,
    }


def pointcloud_project_fast(cfg, point_cloud, transform, predicted_translation, all_rgb,
                            kernel=None, scaling_factor=None, focal_length=None):
    """The whole projection (point_cloud_to.py:191-263): returns the reference's
    dict {proj [P,V,V,1], voxels [P,Vz,V,V,1], tr_pc [P,N,3], voxels_rgb,
    proj_rgb, drc_probs [Vz+1,P,V,V,1], proj_depth [P,V,V,1]}.

    The blur always runs when ``kernel`` is given (the reference's CUDA branch,
    :207-209); ``kernel=None`` means no blur with the TF original's layout."""
    _check_quaternion_cfg(cfg)
    if getattr(cfg, "ptn_max_projection", False):
        raise NotImplementedError("ptn_max_projection is broken in the reference "
                                  "(point_cloud_to.py:234,242) and not supported")
    pts = ops._f32(point_cloud, "point_cloud")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError("point_cloud must be [P,N,3], got %s" % (tuple(pts.shape),))
    P, N, _ = pts.shape
    quat = ops._f32(transform, "transform", (P, 4))
    trans = ops._f32(predicted_translation, "predicted_translation", (P, 3))
    scale = _vec(scaling_factor, "scaling_factor", P)
    focal = _vec(focal_length, "focal_length", P)
    col = None if all_rgb is None else _features(all_rgb, pts)
    params = ops.make_params(cfg, P, N, flip_y=True)
    taps = ops.host_taps(kernel)
    # the colour integral needs the ray-event probabilities, asked for or not
    mask, depth, tr_pc, voxels, probs = ops.project(
        pts, quat, trans, focal, scale, params, taps,
        _options["voxels"], _options["drc_probs"] or col is not None, _scatter_mode(),
        _options["plane_local"])
    voxels_rgb = proj_rgb = None
    if col is not None:
        voxels_rgb, proj_rgb = _project_features(cfg, tr_pc, col, probs, taps, P, N)
    return {
        "proj": mask.unsqueeze(-1),
        "voxels": None if voxels is None else voxels.unsqueeze(-1),
        "tr_pc": tr_pc,
        "voxels_rgb": voxels_rgb,
        "proj_rgb": proj_rgb,
        "drc_probs": None if (probs is None or not _options["drc_probs"]) else probs.unsqueeze(-1),
        "proj_depth": depth.unsqueeze(-1),
    }


def _project_features(cfg, tr_pc, col, probs, taps, P, N):
    """The rgb branch of the TF original's pointcloud_project_fast (point_cloud.py:244-277; the
    torch port's branch is broken, point_cloud_to.py:64): feature scatter with the occupancy's
    weights, clip + per-channel blur, optional division by the blurred raw occupancy, optional
    clip after the blur, Y flip and colour integral with a white background.
    Returns (voxels_rgb [P,Vz,V,V,C] -- not differentiable --, proj_rgb [P,V,V,C])."""
    clip_after = bool(getattr(cfg, "pc_rgb_clip_after_conv", False))
    params = ops.make_params(cfg, P, N, flip_y=True)
    tr = tr_pc.detach() if getattr(cfg, "pc_rgb_stop_points_gradient", False) else tr_pc
    fgrid = ops.FeatGridFn.apply(tr, col, params, taps, not clip_after)
    div, eps = None, 0.0
    if getattr(cfg, "pc_rgb_divide_by_occupancies", False):
        if taps is None:
            raise ValueError("pc_rgb_divide_by_occupancies needs a smoothing kernel "
                             "(point_cloud.py:258 blurs the raw occupancy)")
        with torch.no_grad():       # stop_gradient(voxels_raw), point_cloud.py:257
            raw = ops.ScatterFn.apply(tr_pc.detach(), params, _scatter_mode())
            div = ops._blur3d(raw, params, taps)
        eps = float(getattr(cfg, "pc_rgb_divide_by_occupancies_epsilon", 0.01))
    proj_rgb, voxels_rgb = ops.ColourFn.apply(probs, fgrid, div, eps, clip_after, params,
                                              _options["voxels"])
    return voxels_rgb, proj_rgb
