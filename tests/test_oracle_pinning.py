"""CPU: the oracle against the reference's golden vectors and (when the
reference tree is mounted) against the reference executed live."""
import numpy as np
import pytest
import torch

import _golden
import _inputs
from oracle import closed_form as CF
from oracle import ref_loader as RL
from oracle.config import default_cfg

CASE_NAMES = sorted(_golden.CASES)
# the two heaviest cases take a few seconds each on CPU; keep them, they pin
# the shapes that matter (config 1 / config 4)


def _run_oracle(name, rec):
    cfg = _golden.case_cfg(name)
    inp = _golden.case_inputs(rec, requires_grad=True)
    out = CF.project(cfg, inp["points"], inp["quat"], inp["translation"], inp["kernel"],
                     inp["scale"], inp["focal"])
    P, V = inp["points"].shape[0], cfg.vox_size
    Wp, Wd = _inputs.loss_weights(P, V)
    loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
    keys = [k for k in ("points", "quat", "translation", "focal", "scale") if inp[k] is not None]
    grads = dict(zip(keys, torch.autograd.grad(loss, [inp[k] for k in keys])))
    return out, loss, grads


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_golden(name):
    rec = _golden.load(name)
    out, loss, grads = _run_oracle(name, rec)
    # forward: the restatement is bit-identical to the reference in fp64
    for k in ("proj", "proj_depth", "tr_pc"):
        assert _golden.rel_err(out[k], rec[k]) < 1e-13, k
    for k in ("voxels", "drc_probs", "voxels_raw"):
        flat = out[k].detach().reshape(-1)
        assert abs(flat.sum().item() - rec[k + "_sum"]) <= 1e-9 * abs(rec[k + "_sum"]) + 1e-12
        sub = flat[::_golden.VOX_STRIDE].float()
        assert _golden.rel_err(sub, rec[k + "_sub"]) < 1e-6, k
    assert abs(loss.item() - rec["loss"]) < 1e-9 * abs(rec["loss"])
    # gradients come back fp32 (leaf dtype); autograd accumulation order may differ
    for k, g in grads.items():
        assert _golden.rel_err(g, rec["grad_" + k]) < 2e-6, k


def test_recipe_kat():
    """run/pc_full_proj_test.py:48-61 -- numpy seed 0, sums printed by the reference."""
    rec = _golden.load("recipe_kat")
    cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=21)
    np.random.seed(0)
    cam = torch.from_numpy(np.random.random((128, 4))).float()
    pc = torch.from_numpy(np.random.random((128, 140, 3))).float()
    sf = torch.from_numpy(np.random.random((128, 1))).float()
    lit = CF.project(cfg, pc, cam, None, None, sf)          # reference-on-CPU skips the blur
    blur = CF.project(cfg, pc, cam, None, CF.smoothing_taps(cfg, 3.0), sf)
    for tag, out in (("literal", lit), ("blurred", blur)):
        for k in ("proj", "voxels", "tr_pc", "drc_probs", "proj_depth"):
            want = float(rec["%s_%s_sum" % (tag, k)])
            assert abs(out[k].sum().item() - want) <= 1e-10 * abs(want), (tag, k)
    assert _golden.rel_err(blur["proj"][:8], rec["blurred_proj_first8"]) < 1e-13
    assert _golden.rel_err(blur["proj_depth"][:8], rec["blurred_proj_depth_first8"]) < 1e-13


def test_golden_taps_match_host_mirror():
    import pytorch_unsup_pc_b200 as dpc
    for name in CASE_NAMES:
        spec, rec = _golden.CASES[name], _golden.load(name)
        if spec["sigma"] is None:
            continue
        k = dpc.smoothing_kernel(_golden.case_cfg(name), spec["sigma"])
        o = CF.smoothing_taps(_golden.case_cfg(name), spec["sigma"])
        for got, orc, key in zip(k, o, ("taps_x", "taps_y", "taps_z")):
            assert np.array_equal(got.reshape(-1).numpy(), rec[key]), (name, key)
            assert torch.equal(got, orc)


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    cfg = default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    case = _inputs.make_case(cfg, 3, 500, seed=77, translation=True, focal=True, screened=False)
    kern = RL.ref_smoothing_kernel(cfg, 1.2)
    ref = RL.ref_project(cfg, case["points"], case["quat"], case["translation"], kern,
                         case["scale"], case["focal"])
    orc = CF.project(cfg, case["points"], case["quat"], case["translation"],
                     CF.smoothing_taps(cfg, 1.2), case["scale"], case["focal"])
    for k in ("proj", "proj_depth", "tr_pc", "voxels", "drc_probs", "voxels_raw"):
        assert torch.equal(ref[k], orc[k]), k


MIRRORED = {
    "pc_to": ["pointcloud_project_fast", "pc_perspective_transform", "pointcloud2voxels3d_fast",
              "smoothen_voxels3d", "convolve_rgb", "pc_point_dropout"],
    "drc": ["drc_projection", "drc_depth_projection", "drc_event_probabilities",
            "project_volume_rgb_integral"],
    "gauss_kernel": ["smoothing_kernel", "gauss_kernel_1d", "separable_kernels"],
}


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
def test_mirror_signatures_match_live_reference():
    """The drop-in claim, guarded: every mirrored function takes the reference's parameters, by
    the same names, in the same order, with the same defaults.  A mirror may APPEND optional
    parameters (pc_point_dropout's seed / indices); it may not rename, reorder or drop any."""
    import inspect
    import sys
    import pytorch_unsup_pc_b200 as dpc
    mods = RL.load()
    if RL._DPC not in sys.path:
        sys.path.insert(0, RL._DPC)
    import util.point_cloud_distance as pcd
    mods = dict(mods, pcd=pcd)
    names = dict(MIRRORED, pcd=["point_cloud_distance"])
    for mod, fns in names.items():
        for fn in fns:
            ref = list(inspect.signature(getattr(mods[mod], fn)).parameters.values())
            mine = list(inspect.signature(getattr(dpc, fn)).parameters.values())
            assert len(mine) >= len(ref), fn
            for r, m in zip(ref, mine):
                assert (r.name, r.kind) == (m.name, m.kind), (fn, r, m)
                assert r.default == m.default or (r.default is inspect.Parameter.empty
                                                  and m.default is inspect.Parameter.empty), (fn, r, m)
            for extra in mine[len(ref):]:
                assert extra.default is not inspect.Parameter.empty, (fn, extra)
