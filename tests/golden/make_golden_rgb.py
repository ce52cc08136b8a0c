#!/usr/bin/env python
"""Generate tests/golden/rgb.npz from oracle/rgb.py:    python tests/golden/make_golden_rgb.py

Made by the ORACLE, not by the reference (its torch port of the rgb branch does not run:
util/point_cloud_to.py:64, util/drc.py:137): a regression fixture for the oracle's restatement
and the CUDA path.  The reference-made twin is rgb_tf.npz (make_golden_rgb_tf.py: the
reference's TF original executed through oracle/tf_shim.py);
tests/test_rgb.py::test_oracle_made_fixture_equals_reference_made_fixture compares the two entry
by entry.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import closed_form as CF             # noqa: E402
from oracle import rgb as ORGB                   # noqa: E402
from oracle.config import default_cfg            # noqa: E402
import _inputs                                   # noqa: E402

CASES = {
    "default": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=1.5, P=3, N=500, seed=1201),
    "clip_after_divide": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11, pc_rgb_clip_after_conv=True,
                                       pc_rgb_divide_by_occupancies=True), sigma=1.5, P=2, N=500,
                              seed=1202),
    "stop_grad_noblur": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11,
                                      pc_rgb_stop_points_gradient=True), sigma=None, P=2, N=400,
                             seed=1203),
}


def make_inputs(spec):
    cfg = default_cfg(**spec["cfg"])
    case = _inputs.make_case(cfg, spec["P"], spec["N"], spec["seed"], kind="clustered", scale=True,
                             screened=True)
    g = torch.Generator().manual_seed(spec["seed"] + 1)
    rgb = torch.rand(spec["P"], spec["N"], 3, generator=g)
    W = torch.rand(spec["P"], cfg.vox_size, cfg.vox_size, 3, generator=g)
    kern = None if spec["sigma"] is None else CF.smoothing_taps(cfg, spec["sigma"])
    return cfg, case, rgb, W, kern


def run(spec):
    cfg, case, rgb, W, kern = make_inputs(spec)
    leaves = {k: case[k].clone().requires_grad_() for k in ("points", "quat", "scale")}
    leaves["rgb"] = rgb.clone().requires_grad_()
    out = ORGB.project_rgb(cfg, leaves["points"], leaves["quat"], leaves["rgb"], None, kern,
                           leaves["scale"])
    Wp, Wd = _inputs.loss_weights(spec["P"], cfg.vox_size)
    loss = ((out["proj_rgb"] * W.double()).sum() + (out["proj"] * Wp.double()).sum()
            + 0.1 * (out["proj_depth"] * Wd.double()).sum())
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return out, loss, dict(zip(leaves, grads))


def main():
    rec = {}
    for name, spec in CASES.items():
        out, loss, grads = run(spec)
        rec[name + "/loss"] = np.float64(loss.item())
        rec[name + "/proj_rgb"] = out["proj_rgb"].detach().numpy()
        rec[name + "/voxels_rgb_sum"] = np.float64(out["voxels_rgb"].sum().item())
        rec[name + "/voxels_rgb_sub"] = out["voxels_rgb"].detach().reshape(-1)[::61].numpy().astype(np.float32)
        for k, g in grads.items():
            rec[name + "/grad_" + k] = g.numpy()
        print(name, "loss=%.9f" % loss.item())
    path = os.path.join(HERE, "rgb.npz")
    np.savez_compressed(path, **rec)
    print(os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
