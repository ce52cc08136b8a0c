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
