#!/usr/bin/env python
"""Generate tests/golden/candidate_loss.npz by executing the REAL reference's
``ModelPointCloud.add_proj_loss`` (models/model_pc_to.py:339-385, 410-440) on the
CPU (needs /root/reference):    python tests/golden/make_golden_loss.py

Cases: chair_unsupervised shapes (128^2 masks pooled to 64^2, 4 candidates), the
same with per-sample ``valid_samples`` weights (``variable_num_views``), and an
un-pooled 32^2 case with 3 candidates.  ``pose_predictor_student`` is switched
off: the student loss is outside the row.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader as RL   # noqa: E402

CASES = {
    "pool128_c4": dict(BV=8, C=4, V=64, G=128, weights=False, scale=1.0, seed=41),
    "pool128_c4_weights": dict(BV=6, C=4, V=64, G=128, weights=True, scale=0.5, seed=42),
    "nopool32_c3": dict(BV=5, C=3, V=32, G=32, weights=False, scale=2.0, seed=43),
}


def make_inputs(spec):
    g = torch.Generator().manual_seed(spec["seed"])
    BV, C, V, G = spec["BV"], spec["C"], spec["V"], spec["G"]
    masks = (torch.rand(BV, 1, G, G, generator=g) > 0.5).float()
    projs = torch.rand(BV * C, V, V, 1, generator=g)
    w = (torch.rand(BV, generator=g) > 0.3).float() + 0.25 if spec["weights"] else None
    return masks, projs, w


def main():
    out = {}
    for name, spec in CASES.items():
        cfg = RL.reference_cfg(pose_predict_num_candidates=spec["C"], pose_predictor_student=False,
                               variable_num_views=spec["weights"])
        masks, projs, w = make_inputs(spec)
        p = projs.double().requires_grad_()
        total, min_loss = RL.ref_candidate_loss(cfg, masks.clone(), p, spec["scale"], w)
        (gp,) = torch.autograd.grad(total, p)
        out[name + "/masks"] = masks.numpy()
        out[name + "/projs"] = projs.numpy()
        if w is not None:
            out[name + "/weights"] = w.numpy()
        out[name + "/total"] = np.float64(total.item())
        out[name + "/min_loss"] = min_loss.numpy()
        out[name + "/g_projs"] = gp.numpy()
        print(name, total.item(), min_loss.tolist())
    np.savez_compressed(os.path.join(HERE, "candidate_loss.npz"), **out)


if __name__ == "__main__":
    main()
