#!/usr/bin/env python
"""Generate tests/golden/render_loss.npz by executing the REAL reference (needs /root/reference):
    python tests/golden/make_golden_render_loss.py

The step between the decoder and ``loss.backward()`` of ``ModelPointCloud``: the reference's own
``tf_repeat_0`` + ``pc_point_dropout`` + projection (models/model_pc_to.py:302-331, composed in
its CUDA-branch order by oracle/ref_loader.py) followed by its own ``add_proj_loss`` /
``proj_loss_pose_candidates`` (:339-385, 410-440), and torch autograd through all of it.
Inputs are boundary-screened against every replica's pose (tests/_inputs.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader as RL              # noqa: E402
from golden.make_golden_replicas import make_inputs   # noqa: E402

CASES = {
    # 3 clouds x 2 views x 2 candidates, 64^2 masks pooled to 32^2
    "v32_pool": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=1.5, B=3, views=2, cands=2,
                     N=600, keep=1, G=64, weights=False, wscale=1.0, seed=2101),
    # dropout (70 % kept), 3 candidates, per-view weights, un-pooled masks, weight_scale 0.5
    "v32_drop_weights": dict(cfg=dict(vox_size=32, pc_gauss_kernel_size=11), sigma=1.0, B=2, views=2,
                             cands=3, N=700, keep=0.7, G=32, weights=True, wscale=0.5, seed=2102),
    # chair_unsupervised shapes: 64^3, K=21, sigma 3, 128^2 masks, 4 candidates
    "v64_chair": dict(cfg=dict(vox_size=64, pc_gauss_kernel_size=21), sigma=3.0, B=2, views=2, cands=4,
                      N=2000, keep=1, G=128, weights=False, wscale=1.0, seed=2103),
}


def make_masks(spec, cfg, pts, quat, scale, kernel):
    """Ground truth with structure: the silhouette of the same clouds under perturbed poses,
    upsampled to G x G and binarised -- so that candidates differ in loss the way they do in
    training -- plus per-view weights."""
    g = torch.Generator().manual_seed(spec["seed"] + 13)
    BV = spec["B"] * spec["views"]
    C = spec["cands"]
    pick = torch.randint(0, C, (BV,), generator=g)                       # the candidate nearest the truth
    q_gt = quat.reshape(BV, C, 4)[torch.arange(BV), pick] + 0.15 * torch.randn(BV, 4, generator=g)
    from oracle import closed_form as CF
    from oracle.replicas import tf_repeat_0
    out = CF.project(cfg, tf_repeat_0(pts, spec["views"]), q_gt, None, kernel,
                     scale.reshape(BV, C, 1)[:, 0])
    sil = out["proj"].float().permute(0, 3, 1, 2)                      # [BV,1,V,V]
    up = torch.nn.functional.interpolate(sil, size=(spec["G"], spec["G"]), mode="bilinear",
                                         align_corners=False)
    masks = (up > 0.35).float()
    w = ((torch.rand(BV, generator=g) > 0.3).float() + 0.25) if spec["weights"] else None
    return masks, w


def main():
    rec = {}
    for name, spec in CASES.items():
        cfg, pts, quat, scale = make_inputs(spec)
        kernel = RL.ref_smoothing_kernel(cfg, spec["sigma"])
        masks, w = make_masks(spec, cfg, pts, quat, scale, kernel)
        rcfg = RL.reference_cfg(pose_predict_num_candidates=spec["cands"], pose_predictor_student=False,
                                variable_num_views=spec["weights"], **spec["cfg"])
        leaves = [t.clone().requires_grad_() for t in (pts, quat, scale)]
        out, idx = RL.ref_project_replicated(rcfg, leaves[0], leaves[1], spec["views"], spec["cands"],
                                             spec["keep"], spec["seed"], None, kernel, leaves[2])
        total, min_loss = RL.ref_candidate_loss(rcfg, masks.clone(), out["proj"], spec["wscale"], w)
        grads = torch.autograd.grad(total, leaves)
        rec[name + "/in_points"] = pts.numpy()
        rec[name + "/in_quat"] = quat.numpy()
        rec[name + "/in_scale"] = scale.numpy()
        rec[name + "/in_masks"] = masks.numpy().astype(np.uint8)
        if w is not None:
            rec[name + "/in_weights"] = w.numpy()
        for i, k in enumerate("xyz"):
            rec[name + "/taps_" + k] = kernel[i].reshape(-1).numpy()
        if idx is not None:
            rec[name + "/indices"] = idx.numpy().astype(np.int32)
        rec[name + "/loss"] = np.float64(total.item())
        rec[name + "/min_loss"] = min_loss.numpy()
        rec[name + "/proj"] = out["proj"].detach().float().numpy()
        for k, g in zip(("points", "quat", "scale"), grads):
            rec[name + "/grad_" + k] = g.numpy()
        print(name, "loss=%.9f" % total.item(), "min_loss", min_loss.tolist())
    path = os.path.join(HERE, "render_loss.npz")
    np.savez_compressed(path, **rec)
    print(os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
