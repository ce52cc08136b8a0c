"""Helpers shared by the golden-vector tests."""
import os

import numpy as np
import torch

from golden.cases import CASES, VOX_STRIDE  # noqa: F401
from oracle.config import default_cfg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def case_cfg(name):
    return default_cfg(**CASES[name]["cfg"])


def case_inputs(rec, device="cpu", requires_grad=False):
    """Tensors for (points, quat, translation, focal, scale) + kernel list."""
    out = {}
    for k in ("points", "quat", "translation", "focal", "scale"):
        if "in_" + k in rec:
            t = torch.from_numpy(rec["in_" + k]).to(device)
            out[k] = t.requires_grad_() if requires_grad else t
        else:
            out[k] = None
    if "taps_x" in rec:
        kx, ky, kz = (torch.from_numpy(rec[n]) for n in ("taps_x", "taps_y", "taps_z"))
        out["kernel"] = [kx.reshape(1, 1, 1, 1, -1), ky.reshape(1, 1, 1, -1, 1),
                         kz.reshape(1, 1, -1, 1, 1)]
    else:
        out["kernel"] = None
    return out


def rel_err(a, b):
    """max|a-b| / max|b| -- the scale-relative error of SURVEY.md section 7 (hard part 2)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
