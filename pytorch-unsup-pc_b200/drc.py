"""Drop-in mirror of the reference's util/drc.py (the functions on the path)."""
from . import ops


def _grid(voxels, cfg):
    vox = ops._f32(voxels, "voxels")
    if vox.dim() != 5 or vox.shape[-1] != 1:
        raise ValueError("voxels must be [P,Z,Y,X,1], got %s" % (tuple(vox.shape),))
    P, Z, Y, X, _ = vox.shape
    params = ops.make_params(cfg, P, 0, flip_y=False)
    if Y != X or X != params.V:
        raise ValueError("voxels %s does not match cfg.vox_size %d" % (tuple(vox.shape), params.V))
    params.Vz = Z
    return vox.reshape(P, Z, Y, X), params


def _check(cfg):
    if not getattr(cfg, "drc_tf_cumulative", True):
        raise NotImplementedError("drc_tf_cumulative: false is broken in the reference "
                                  "(drc.py:45, NameError) and not supported")


def drc_event_probabilities(voxels, cfg):
    """voxels [P,Z,Y,X,1] -> p [Z+1,P,Y,X,1] (drc.py:48-111)."""
    _check(cfg)
    vox, params = _grid(voxels, cfg)
    _, probs, _ = ops.DrcFn.apply(vox, params)
    return probs.unsqueeze(-1)


def drc_projection(voxels, cfg):
    """voxels [P,Z,Y,X,1] -> (proj [P,Y,X,1], p [Z+1,P,Y,X,1]) (drc.py:114-129)."""
    _check(cfg)
    vox, params = _grid(voxels, cfg)
    mask, probs, _ = ops.DrcFn.apply(vox, params)
    return mask.unsqueeze(-1), probs.unsqueeze(-1)


def drc_depth_projection(p, cfg):
    """p [Z+1,P,Y,X,1] -> expected depth [P,Y,X,1] (drc.py:152-160)."""
    probs = ops._f32(p, "p")
    if probs.dim() != 5 or probs.shape[-1] != 1:
        raise ValueError("p must be [Z+1,P,Y,X,1], got %s" % (tuple(probs.shape),))
    Z1, P, Y, X, _ = probs.shape
    params = ops.make_params(cfg, P, 0, flip_y=False)
    if Y != X or X != params.V:
        raise ValueError("p %s does not match cfg.vox_size %d" % (tuple(probs.shape), params.V))
    params.Vz = Z1 - 1
    return ops.DepthFromProbsFn.apply(probs.reshape(Z1, P, Y, X), params).unsqueeze(-1)


def project_volume_rgb_integral(cfg, p, rgb):
    """p [Z+1,P,Y,X,1], rgb [P,Z,Y,X,C] -> [P,Y,X,C]: sum_k p_k rgb_k with a white background as
    the last ray event (drc.py:132-142; the torch port's version does not run -- torch.float63
    at :137 -- so this follows the TF original's util/drc.py of the same name)."""
    probs = ops._f32(p, "p")
    col = ops._f32(rgb, "rgb")
    if probs.dim() != 5 or probs.shape[-1] != 1:
        raise ValueError("p must be [Z+1,P,Y,X,1], got %s" % (tuple(probs.shape),))
    Z1, P, Y, X, _ = probs.shape
    if col.dim() != 5 or tuple(col.shape[:4]) != (P, Z1 - 1, Y, X) or not 1 <= col.shape[4] <= 4:
        raise ValueError("rgb must be [P,Z,Y,X,C<=4] matching p %s, got %s"
                         % (tuple(probs.shape), tuple(col.shape)))
    params = ops.make_params(cfg, P, 0, flip_y=False)       # both inputs are already flipped
    if Y != X or X != params.V:
        raise ValueError("p %s does not match cfg.vox_size %d" % (tuple(probs.shape), params.V))
    params.Vz = Z1 - 1
    planar = col.permute(0, 4, 1, 2, 3).contiguous()
    proj, _ = ops.ColourFn.apply(probs.reshape(Z1, P, Y, X), planar, None, 0.0, False, params, False)
    return proj
