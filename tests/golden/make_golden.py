#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REAL reference on the CPU.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

For every case in cases.py the seeded, boundary-screened inputs are pushed
through the reference's own functions (oracle.ref_loader.ref_project: the
CUDA-branch composition of util/point_cloud_to.py:191-263) and the loss
``sum(proj*Wp) + 0.1*sum(depth*Wd)`` is back-propagated with the reference's
own autograd graph.  Inputs, outputs and gradients are stored; the two large
tensors (voxels, drc_probs) are stored as a strided subsample plus sums.
The recipe KAT reproduces run/pc_full_proj_test.py:48-61 (numpy seed 0).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader as RL              # noqa: E402
from oracle.config import default_cfg            # noqa: E402
from golden.cases import CASES, RECIPE, VOX_STRIDE  # noqa: E402
import _inputs                                   # noqa: E402


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def ref_aniso_kernel(cfg, sigma):
    """What gauss_kernel.py:38-51 intends (its own :49 reshape crashes)."""
    gk = RL.load()["gauss_kernel"]
    fsz = cfg.pc_gauss_kernel_size
    ratio = cfg.vox_size_z / cfg.vox_size
    fsz_z = int(np.floor(fsz * ratio))
    if fsz_z % 2 == 0:
        fsz_z += 1
    k = gk.gauss_kernel_1d(fsz, sigma)
    kz = gk.gauss_kernel_1d(fsz_z, sigma * ratio)
    return [k.reshape(1, 1, 1, 1, fsz), k.reshape(1, 1, 1, fsz, 1), kz.reshape(1, 1, fsz_z, 1, 1)]


def run_case(name, spec):
    cfg = default_cfg(**spec["cfg"])
    case = _inputs.make_case(cfg, spec["P"], spec["N"], spec["seed"],
                             kind=spec.get("kind", "uniform"),
                             translation=spec.get("translation", False),
                             focal=spec.get("focal", False),
                             scale=spec.get("scale", True), screened=True)
    if spec["sigma"] is None:
        kernel = None
    elif spec.get("aniso_kernel"):
        kernel = ref_aniso_kernel(cfg, spec["sigma"])
    else:
        kernel = RL.ref_smoothing_kernel(cfg, spec["sigma"])
    leaves = {}
    for k in ("points", "quat", "translation", "focal", "scale"):
        leaves[k] = None if case[k] is None else case[k].clone().requires_grad_()
    out = RL.ref_project(cfg, leaves["points"], leaves["quat"], leaves["translation"], kernel,
                         leaves["scale"], leaves["focal"])
    V = cfg.vox_size
    Wp, Wd = _inputs.loss_weights(spec["P"], V)
    loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
    keys = [k for k in leaves if leaves[k] is not None]
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys])
    rec = {"loss": np.float64(loss.item())}
    for k in case:
        if case[k] is not None:
            rec["in_" + k] = _np(case[k])
    if kernel is not None:
        rec["taps_x"] = _np(kernel[0].reshape(-1))
        rec["taps_y"] = _np(kernel[1].reshape(-1))
        rec["taps_z"] = _np(kernel[2].reshape(-1))
    rec["proj"] = _np(out["proj"])
    rec["proj_depth"] = _np(out["proj_depth"])
    rec["tr_pc"] = _np(out["tr_pc"])
    for big in ("voxels", "drc_probs", "voxels_raw"):
        flat = out[big].detach().reshape(-1)
        rec[big + "_sub"] = _np(flat[::VOX_STRIDE]).astype(np.float32)
        rec[big + "_sum"] = np.float64(flat.sum().item())
        rec[big + "_sqsum"] = np.float64((flat * flat).sum().item())
    for k, g in zip(keys, grads):
        rec["grad_" + k] = _np(g)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print("%-24s loss=%.9f  %d KB" % (name, loss.item(), os.path.getsize(path) // 1024))


def run_recipe():
    cfg = default_cfg(**RECIPE["cfg"])
    np.random.seed(0)
    cam = torch.from_numpy(np.random.random((128, 4))).float()
    pc = torch.from_numpy(np.random.random((128, 140, 3))).float()
    sf = torch.from_numpy(np.random.random((128, 1))).float()
    kernel = RL.ref_smoothing_kernel(cfg, RECIPE["sigma"])
    rec = {}
    # (a) the reference function called literally: on a CPU-only box it skips
    # the blur (point_cloud_to.py:206-212) -- these are the five sums the
    # reference's own script prints
    lit = RL.ref_project_literal(cfg, pc, cam, None, kernel, sf)
    # (b) the CUDA-branch composition (blur on)
    full = RL.ref_project(cfg, pc, cam, None, kernel, sf)
    for tag, out in (("literal", lit), ("blurred", full)):
        for k in ("proj", "voxels", "tr_pc", "drc_probs", "proj_depth"):
            rec["%s_%s_sum" % (tag, k)] = np.float64(out[k].sum().item())
    rec["blurred_proj_first8"] = _np(full["proj"][:8])
    rec["blurred_proj_depth_first8"] = _np(full["proj_depth"][:8])
    path = os.path.join(HERE, "recipe_kat.npz")
    np.savez_compressed(path, **rec)
    print("recipe_kat", {k: float(v) for k, v in rec.items() if k.endswith("_sum")})


if __name__ == "__main__":
    torch.manual_seed(0)
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(name, spec)
    if not only or "recipe_kat" in only:
        run_recipe()
