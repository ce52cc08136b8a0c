#!/usr/bin/env python
"""Generate tests/golden/replicas.npz by executing the REAL reference on the CPU
(needs /root/reference):    python tests/golden/make_golden_replicas.py

Next-row f2: the reference's own ``tf_repeat_0`` (models/model_pc_to.py:47-56), its own
``pc_point_dropout`` (util/point_cloud_to.py:269-295, numpy global RNG seeded) and its
projection functions, with the loss ``sum(proj*Wp) + 0.1*sum(depth*Wd)`` back-propagated by
the reference's autograd graph to the UN-replicated cloud tensor.  Two cases: replication
only, and replication + dropout (the sampled indices are stored with the outputs).
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
import _inputs                                   # noqa: E402

CASES = {
    # 3 clouds x 2 views x 2 candidates, nothing dropped
    "rep4": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=1.5, B=3, views=2, cands=2,
                 N=600, keep=1, seed=1101),
    # 2 clouds x 4 views x 1 candidate... x 3 candidates, 7 of 10 points kept
    "rep6_drop": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=1.5, B=2, views=2,
                      cands=3, N=700, keep=0.7, seed=1102),
}


def make_inputs(spec):
    cfg = default_cfg(**spec["cfg"])
    P = spec["B"] * spec["views"] * spec["cands"]
    case = _inputs.make_case(cfg, P, spec["N"], spec["seed"], translation=False, focal=False,
                             scale=True, screened=False)
    pts = case["points"][:spec["B"]].contiguous()
    # boundary screening (tests/_inputs.py) of every cloud against ALL of its replicas' poses
    from oracle import closed_form as CF
    from oracle.replicas import tf_repeat_0
    R = spec["views"] * spec["cands"]
    g = torch.Generator().manual_seed(spec["seed"] + 7)
    for _ in range(50):
        tr = CF.pose_transform(cfg, tf_repeat_0(pts, R), case["quat"], None, None)
        bad = _inputs.near_boundary(cfg, tr).reshape(spec["B"], R, -1).any(dim=1)
        if int(bad.sum()) == 0:
            break
        pts[bad] = (torch.rand(int(bad.sum()), 3, generator=g) - 0.5) * 0.9
    else:
        raise RuntimeError("screening did not converge")
    return cfg, pts, case["quat"], case["scale"]


def main():
    rec = {}
    for name, spec in CASES.items():
        cfg, pts, quat, scale = make_inputs(spec)
        P = quat.shape[0]
        kernel = RL.ref_smoothing_kernel(cfg, spec["sigma"])
        leaves = [t.clone().requires_grad_() for t in (pts, quat, scale)]
        out, idx = RL.ref_project_replicated(cfg, leaves[0], leaves[1], spec["views"], spec["cands"],
                                             spec["keep"], spec["seed"], None, kernel, leaves[2])
        Wp, Wd = _inputs.loss_weights(P, cfg.vox_size)
        loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
        grads = torch.autograd.grad(loss, leaves)
        rec[name + "/in_points"] = pts.numpy()
        rec[name + "/in_quat"] = quat.numpy()
        rec[name + "/in_scale"] = scale.numpy()
        for i, k in enumerate("xyz"):
            rec[name + "/taps_" + k] = kernel[i].reshape(-1).numpy()
        if idx is not None:
            rec[name + "/indices"] = idx.numpy().astype(np.int32)
        rec[name + "/loss"] = np.float64(loss.item())
        for k in ("proj", "proj_depth", "tr_pc"):
            rec[name + "/" + k] = out[k].detach().numpy()
        for k, g in zip(("points", "quat", "scale"), grads):
            rec[name + "/grad_" + k] = g.numpy()
        print(name, "loss=%.9f" % loss.item(), "tr_pc", tuple(out["tr_pc"].shape))
    path = os.path.join(HERE, "replicas.npz")
    np.savez_compressed(path, **rec)
    print(os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
