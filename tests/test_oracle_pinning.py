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


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
def test_host_taps_equal_live_reference_over_the_sigma_schedule():
    """The taps are an input of the path: the product's host mirror (and the oracle's) must hand
    the kernels the reference's exact fp32 weights at every sigma of the 3.0 -> 0.2 schedule
    (model_pc_to.py:59-63), for every odd kernel size up to 21."""
    import pytorch_unsup_pc_b200 as dpc
    for K in (1, 3, 5, 11, 15, 21):
        cfg = default_cfg(vox_size=64, pc_gauss_kernel_size=K)
        for step in range(0, 61):
            sigma = 3.0 + (0.2 - 3.0) * step / 60.0
            ref = RL.ref_smoothing_kernel(cfg, sigma)
            mine = dpc.smoothing_kernel(cfg, sigma)
            orc = CF.smoothing_taps(cfg, sigma)
            assert len(ref) == len(mine) == len(orc) == 3
            for r, m, o in zip(ref, mine, orc):
                assert r.dtype == m.dtype == torch.float32 and r.shape == m.shape
                assert torch.equal(r, m) and torch.equal(r, o), (K, sigma)
            assert torch.equal(dpc.gauss_kernel_1d(K, sigma), RL.load()["gauss_kernel"].gauss_kernel_1d(K, sigma))


def _sweep_draw(i):
    """Draw i of a seeded sweep over the argument space of the path (the reference runs each)."""
    g = torch.Generator().manual_seed(9100 + i)
    pick = lambda xs: xs[int(torch.randint(len(xs), (1,), generator=g))]   # noqa: E731
    v = pick([16, 32, 32, 64])
    spec = dict(V=v, Vz=pick([-1, -1, v // 2]) if i % 3 == 2 else -1, K=pick([1, 3, 5, 11, 21]),
                sigma=pick([0.2, 0.5, 1.0, 2.0, 3.0]), P=pick([1, 2, 3]), N=pick([1, 17, 300, 900]),
                kind=pick(["uniform", "clustered", "recipe"]), translation=bool(pick([0, 1])),
                focal=bool(pick([0, 1])), scale=bool(pick([0, 1, 1])), blur=bool(pick([0, 1, 1, 1])),
                logsum=bool(pick([0, 1, 1, 1])))
    if spec["Vz"] != -1:
        spec["blur"] = False        # the reference's anisotropic kernel raises (gauss_kernel.py:49)
    return spec


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("i", range(16))
def test_oracle_matches_live_reference_sweep(i):
    """The pin of `closed_form` widened from one configuration to a seeded sweep: grid 16^3 / 32^3 /
    64^3 and anisotropic depth, 1-21 taps over the sigma schedule, clouds of 1-900 points incl.
    out-of-frustum and piled-up ones (unscreened), every optional input on and off, both DRC
    forms.  Forward: every output `torch.equal` to the reference executed live; gradients (the
    reference's own autograd against autograd over the restatement): fp32 rounding of the leaves."""
    sp = _sweep_draw(i)
    cfg = default_cfg(vox_size=sp["V"], vox_size_z=sp["Vz"], pc_gauss_kernel_size=sp["K"],
                      drc_logsum=sp["logsum"])
    case = _inputs.make_case(cfg, sp["P"], sp["N"], 9200 + i, kind=sp["kind"], translation=sp["translation"],
                             focal=sp["focal"], scale=sp["scale"], screened=False)
    # the reference raises on a coordinate of exactly +0.5 (point_cloud_to.py:59); clamped clouds hit it
    if sp["kind"] == "clustered":
        case["points"] = case["points"].clamp(-0.499, 0.499)
    keys = [k for k in ("points", "quat", "translation", "focal", "scale") if case[k] is not None]
    vz = sp["V"] if sp["Vz"] == -1 else sp["Vz"]
    Wp, Wd = _inputs.loss_weights(sp["P"], sp["V"])
    res = []
    for which in ("reference", "oracle"):
        lv = {k: case[k].clone().requires_grad_() for k in keys}
        if which == "reference":
            kern = RL.ref_smoothing_kernel(cfg, sp["sigma"]) if sp["blur"] else None
            out = RL.ref_project(cfg, lv["points"], lv["quat"], lv.get("translation"), kern,
                                 lv.get("scale"), lv.get("focal"))
        else:
            kern = CF.smoothing_taps(cfg, sp["sigma"]) if sp["blur"] else None
            out = CF.project(cfg, lv["points"], lv["quat"], lv.get("translation"), kern,
                             lv.get("scale"), lv.get("focal"))
        assert out["voxels"].shape == (sp["P"], vz, sp["V"], sp["V"], 1)
        loss = (out["proj"] * Wp.double()).sum() + 0.1 * (out["proj_depth"] * Wd.double()).sum()
        res.append((out, torch.autograd.grad(loss, [lv[k] for k in keys], allow_unused=True)))
    (ref, gref), (orc, gorc) = res
    for k in ("proj", "proj_depth", "tr_pc", "voxels", "drc_probs", "voxels_raw"):
        assert torch.equal(ref[k], orc[k]), (sp, k)
    for k, a, b in zip(keys, gref, gorc):
        assert (a is None) == (b is None), (sp, k)
        if a is not None and float(a.abs().max()) > 0:
            assert _golden.rel_err(b, a) < 5e-6, (sp, k)


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
