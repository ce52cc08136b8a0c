"""Load and drive the REAL reference implementation (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` does not
exist on the GPU box; nothing that runs there may call into this module
(``available()`` is the guard).  It is used to

* pin ``oracle.closed_form`` against the reference executed live
  (tests/test_oracle_pinning.py), and
* generate the committed golden vectors (tests/golden/make_golden.py).

The reference cannot be imported unmodified under numpy 2.x:
util/quaternion.py:22-23 evaluates ``np.maximum_sctype(np.float)`` at import
time.  A two-line shim restores those names; no reference file is copied or
edited.  The reference's own config loader (util/config.py:46,131-155) needs
easydict and an unsafe ``yaml.load``; we hand the functions a plain attribute
dict with the same keys instead.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

from .config import AttrDict

REFERENCE_ROOT = os.environ.get("DPC_REFERENCE_ROOT", "/root/reference")
_DPC = os.path.join(REFERENCE_ROOT, "dpc")

_mods = None


def available():
    return os.path.isfile(os.path.join(_DPC, "util", "point_cloud_to.py"))


def load():
    """Import the reference's hot-path modules; returns a namespace dict."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference tree not present at %s" % _DPC)
    # numpy>=1.24 removed these; util/quaternion.py:22-23 needs them at import
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "maximum_sctype"):
        np.maximum_sctype = lambda t: np.float64
    if _DPC not in sys.path:
        sys.path.insert(0, _DPC)
    import util.point_cloud_to as pc_to
    import util.drc as drc
    import util.gauss_kernel as gk
    import util.quaternion as quat
    _mods = dict(pc_to=pc_to, drc=drc, gauss_kernel=gk, quaternion=quat)
    return _mods


def reference_cfg(experiment="chair_unsupervised", **overrides):
    """default_config.yaml overlaid with experiments/<experiment>/config.yaml."""
    import yaml
    with open(os.path.join(_DPC, "resources", "default_config.yaml")) as f:
        cfg = AttrDict(yaml.safe_load(f))
    if experiment:
        p = os.path.join(REFERENCE_ROOT, "experiments", experiment, "config.yaml")
        with open(p) as f:
            cfg.update(yaml.safe_load(f))
    cfg.update(overrides)
    return cfg


@contextlib.contextmanager
def _quiet():
    # pointcloud2voxels3d_fast prints "Voxel_time ..." (point_cloud_to.py:85)
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def ref_smoothing_kernel(cfg, sigma):
    return load()["gauss_kernel"].smoothing_kernel(cfg, sigma)


def ref_project(cfg, point_cloud, transform, predicted_translation=None,
                kernel=None, scaling_factor=None, focal_length=None):
    """The reference's projection composed in its CUDA-branch order.

    ``pointcloud_project_fast`` (point_cloud_to.py:191-263) skips the blur
    when ``torch.cuda.is_available()`` is False (:206-212), so on this
    CPU-only container we call the reference's own sub-functions in the order
    the CUDA branch runs them.  ``kernel=None`` follows the TF original
    (point_cloud.py:237-243): no blur, [P,Z,Y,X,1] layout.
    Returns the same dict keys as the reference (fp64 tensors).
    """
    m = load()
    pc_to, drc = m["pc_to"], m["drc"]
    with _quiet():
        tr_pc = pc_to.pc_perspective_transform(cfg, point_cloud, transform,
                                               predicted_translation, focal_length)
        voxels, _ = pc_to.pointcloud2voxels3d_fast(cfg, tr_pc, None)
    voxels = voxels.unsqueeze(1)
    voxels_raw = voxels
    voxels = torch.clamp(voxels, 0.0, 1.0)
    if kernel is not None:
        voxels = pc_to.smoothen_voxels3d(cfg, voxels, kernel)
    voxels = voxels.squeeze(1).unsqueeze(-1)
    if scaling_factor is not None:
        sz = scaling_factor.shape[0]
        voxels = voxels * scaling_factor.reshape(sz, 1, 1, 1, 1)
        voxels = torch.clamp(voxels, 0.0, 1.0)
    proj, drc_probs = drc.drc_projection(voxels, cfg)
    drc_probs = torch.flip(drc_probs, [2])
    proj_depth = drc.drc_depth_projection(drc_probs, cfg)
    proj = torch.flip(proj, [1])
    return {"proj": proj, "voxels": voxels, "tr_pc": tr_pc, "voxels_rgb": None,
            "proj_rgb": None, "drc_probs": drc_probs, "proj_depth": proj_depth,
            "voxels_raw": voxels_raw.squeeze(1)}


def ref_project_literal(cfg, point_cloud, transform, predicted_translation=None,
                        kernel=None, scaling_factor=None, focal_length=None):
    """The reference's ``pointcloud_project_fast`` called as-is (no blur on CPU)."""
    with _quiet():
        return load()["pc_to"].pointcloud_project_fast(
            cfg, point_cloud, transform, predicted_translation, None, kernel,
            scaling_factor=scaling_factor, focal_length=focal_length)


def ref_candidate_loss(cfg, masks, projs, weight_scale=1.0, valid_samples=None):
    """The reference's OWN ``add_proj_loss`` / ``proj_loss_pose_candidates``
    (models/model_pc_to.py:339-385, 410-440) called unbound on a stub ``self``
    (the methods only use ``self.cfg()`` and each other)."""
    load()
    import models.model_pc_to as model_pc

    class _Stub:
        def cfg(self_inner):
            return cfg
        proj_loss_pose_candidates = model_pc.ModelPointCloud.proj_loss_pose_candidates

    inputs = {"masks": masks}
    if valid_samples is not None:
        inputs["valid_samples"] = valid_samples
    outputs = {"projs": projs}
    total, min_loss = model_pc.ModelPointCloud.add_proj_loss(_Stub(), inputs, outputs, weight_scale,
                                                             None, False)
    return total, min_loss


def ref_point_cloud_distance(Vs, Vt):
    """The reference's own util/point_cloud_distance.py:25-40 on the CPU."""
    load()
    import util.point_cloud_distance as pcd
    return pcd.point_cloud_distance(Vs, Vt)


def ref_project_replicated(cfg, point_cloud, transform, step_size, num_candidates, keep_prob,
                           numpy_seed, predicted_translation=None, kernel=None,
                           scaling_factor=None, focal_length=None):
    """The reference's own replication + dropout + projection, in the order of
    ``ModelPointCloud.forward`` / ``compute_projection`` (models/model_pc_to.py:302-306,
    254-265): ``tf_repeat_0`` by views then by candidates (its own function), its own
    ``pc_point_dropout`` with numpy's global RNG seeded to ``numpy_seed``, ``ref_project``.
    Returns (outputs, indices [P,M] int64 recovered by running the same sampler on an
    index-carrying tensor, or None)."""
    m = load()
    import models.model_pc_to as model_pc
    pts = model_pc.tf_repeat_0(point_cloud, step_size)
    if num_candidates > 1:
        pts = model_pc.tf_repeat_0(pts, num_candidates)
    indices = None
    if keep_prob != 1:
        P, N = pts.shape[0], pts.shape[1]
        np.random.seed(numpy_seed)
        carrier = torch.arange(N, dtype=torch.float64).reshape(1, N, 1).repeat(P, 1, 1)
        picked, _ = m["pc_to"].pc_point_dropout(carrier, None, keep_prob)
        indices = picked[:, :, 0].long()
        np.random.seed(numpy_seed)
        pts, _ = m["pc_to"].pc_point_dropout(pts, None, keep_prob)
    out = ref_project(cfg, pts, transform, predicted_translation, kernel, scaling_factor,
                      focal_length)
    return out, indices


# ---------------------------------------------------------------------------------------------
# The point-feature (RGB) branch: the reference's TensorFlow original, executed through
# oracle/tf_shim.py
# ---------------------------------------------------------------------------------------------
_tf_pc = None


def load_tf_point_cloud():
    """Import the reference's TF original ``util/point_cloud.py`` UNMODIFIED with
    ``oracle.tf_shim`` standing in for ``tensorflow``.  What it imports next to TensorFlow
    (``util.drc``, ``util.quaternion``, ``util.camera``, ``util.point_cloud_distance``) are the
    reference's own torch / numpy files, so poses and ray probabilities are computed by exactly
    the code ``ref_project`` runs."""
    global _tf_pc
    if _tf_pc is not None:
        return _tf_pc
    load()
    import importlib
    from . import tf_shim
    had = sys.modules.get("tensorflow")
    sys.modules["tensorflow"] = tf_shim
    try:
        _tf_pc = importlib.import_module("util.point_cloud")
    finally:
        if had is None:
            del sys.modules["tensorflow"]
        else:
            sys.modules["tensorflow"] = had
    return _tf_pc


class _TorchWithTyposRepaired:
    """Stands in for the name ``torch`` inside the reference's util/drc.py while its
    ``project_volume_rgb_integral`` (drc.py:132-142) runs: ``torch.float63`` (:137) reads as
    float64 and ``torch.ones(shape=...)`` (:137) as ``torch.ones(size=...)``.  Everything else is
    torch itself; no reference line is copied or edited."""
    float63 = torch.float64

    @staticmethod
    def ones(*args, shape=None, **kw):
        return torch.ones(*args, **kw) if shape is None else torch.ones(tuple(shape), **kw)

    def __getattr__(self, name):
        return getattr(torch, name)


@contextlib.contextmanager
def _rgb_integral_runnable(drc):
    """util/point_cloud.py:277 hands ``project_volume_rgb_integral`` the TF layout
    [B,Z,Y,X,C]; the function's torch port (the only version in the tree) swaps axes for the
    torch port's channel-first layout [B,C,Z,Y,X] (drc.py:134, fed by point_cloud_to.py:113).
    Hand it that layout and let its own lines run, with the two typos of :137 repaired."""
    orig_fn, orig_torch = drc.project_volume_rgb_integral, drc.torch

    def call(cfg, p, rgb):
        return orig_fn(cfg, p, rgb.permute(0, 4, 1, 2, 3))

    drc.project_volume_rgb_integral, drc.torch = call, _TorchWithTyposRepaired()
    try:
        yield
    finally:
        drc.project_volume_rgb_integral, drc.torch = orig_fn, orig_torch


def tf_kernel(kernel):
    """The reference's separable blur kernels (torch layout [out,in,kd,kh,kw], gauss_kernel.py)
    in the layout ``tf.nn.conv3d`` takes: [kd,kh,kw,in,out]."""
    from . import tf_shim
    if kernel is None:
        return None
    return [tf_shim.wrap(k.permute(2, 3, 4, 1, 0)) for k in kernel]


def ref_project_tf(cfg, point_cloud, transform, predicted_translation=None, all_rgb=None,
                   kernel=None, scaling_factor=None, focal_length=None):
    """The reference's TF-original ``pointcloud_project_fast`` (util/point_cloud.py:229-290)
    called as-is through the shim; returns its output dict as plain torch tensors ([B,Z,Y,X,C]
    grids, [B,Y,X,C] images), still attached to the autograd graph of the arguments."""
    from . import tf_shim
    pc = load_tf_point_cloud()
    drc = load()["drc"]
    w = tf_shim.wrap
    with _rgb_integral_runnable(drc):
        out = pc.pointcloud_project_fast(cfg, w(point_cloud), w(transform), w(predicted_translation),
                                         w(all_rgb), tf_kernel(kernel), w(scaling_factor),
                                         w(focal_length))
    return {k: tf_shim.unwrap(v) for k, v in out.items()}
