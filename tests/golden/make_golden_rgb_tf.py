#!/usr/bin/env python
"""Generate tests/golden/rgb_tf.npz FROM THE REFERENCE:    python tests/golden/make_golden_rgb_tf.py

Build container only (needs /root/reference).  The vectors come from the reference's TensorFlow
original ``util/point_cloud.py`` (``pointcloud_project_fast`` with ``all_rgb``, :229-290) executed
UNMODIFIED through ``oracle/tf_shim.py`` -- a ``tensorflow`` namespace over torch, itself checked
bit for bit against the reference's torch port on the occupancy path
(tests/test_rgb.py::test_tf_shim_reproduces_the_torch_port) -- with the reference's own torch
``util/drc.py`` / ``util/quaternion.py`` underneath, as that file imports them
(``oracle.ref_loader.ref_project_tf``).  Gradients are torch autograd over the reference's lines.

Same cases, inputs, loss and keys as ``make_golden_rgb.py`` (the oracle-made ``rgb.npz``), plus a
64^3 / 21-tap case with translation and focal length, so the two fixtures can be compared entry
by entry."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader as RL              # noqa: E402
from oracle.config import default_cfg            # noqa: E402
import _inputs                                   # noqa: E402
from golden.make_golden_rgb import CASES as BASE_CASES   # noqa: E402

CASES = dict(BASE_CASES)
CASES["chair_all_inputs"] = dict(cfg=dict(vox_size=64, pc_gauss_kernel_size=21), sigma=2.0, P=2, N=2000,
                                 seed=1204, translation=True, focal=True)
INPUT_KEYS = ("points", "quat", "translation", "focal", "scale")


def make_inputs(spec):
    cfg = default_cfg(**spec["cfg"])
    case = _inputs.make_case(cfg, spec["P"], spec["N"], spec["seed"], kind="clustered", scale=True,
                             translation=spec.get("translation", False), focal=spec.get("focal", False),
                             screened=True)
    g = torch.Generator().manual_seed(spec["seed"] + 1)
    rgb = torch.rand(spec["P"], spec["N"], 3, generator=g)
    W = torch.rand(spec["P"], cfg.vox_size, cfg.vox_size, 3, generator=g)
    return cfg, case, rgb, W


def loss_of(out, W, P, V):
    Wp, Wd = _inputs.loss_weights(P, V)
    return ((out["proj_rgb"] * W.double()).sum() + (out["proj"] * Wp.double()).sum()
            + 0.1 * (out["proj_depth"] * Wd.double()).sum())


def run_reference(spec, dtype=None):
    """-> (outputs, loss, gradients by input name) of the reference's TF source.  ``dtype``
    casts the leaves (float64: gradients free of the fp32 leaf rounding)."""
    cfg, case, rgb, W = make_inputs(spec)
    leaves = {k: case[k].clone() for k in INPUT_KEYS if case.get(k) is not None}
    leaves["rgb"] = rgb.clone()
    if dtype is not None:
        leaves = {k: v.to(dtype) for k, v in leaves.items()}
    leaves = {k: v.requires_grad_() for k, v in leaves.items()}
    kern = None if spec["sigma"] is None else RL.ref_smoothing_kernel(cfg, spec["sigma"])
    out = RL.ref_project_tf(cfg, leaves["points"], leaves["quat"], leaves.get("translation"),
                            leaves["rgb"], kern, leaves["scale"], leaves.get("focal"))
    loss = loss_of(out, W, spec["P"], cfg.vox_size)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return out, loss, dict(zip(leaves, grads))


def main():
    assert RL.available(), "the reference tree is needed to make this fixture"
    rec = {}
    for name, spec in CASES.items():
        out, loss, grads = run_reference(spec)
        rec[name + "/loss"] = np.float64(loss.item())
        rec[name + "/proj_rgb"] = out["proj_rgb"].detach().numpy()
        rec[name + "/voxels_rgb_sum"] = np.float64(out["voxels_rgb"].sum().item())
        rec[name + "/voxels_rgb_sub"] = out["voxels_rgb"].detach().reshape(-1)[::61].numpy().astype(np.float32)
        for k, g in grads.items():
            rec[name + "/grad_" + k] = g.numpy()
        print(name, "loss=%.9f" % loss.item())
    path = os.path.join(HERE, "rgb_tf.npz")
    np.savez_compressed(path, **rec)
    print(os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
