"""Seeded synthetic inputs for the parity tests.

Inputs are generated on the CPU with a fixed ``torch.Generator`` so the CPU
oracle and the GPU path see identical bits.  ``screen`` implements the
boundary screening of SURVEY.md section 7 (hard part 1): the gradient of a
trilinear weight is discontinuous at cell faces and the forward is
discontinuous at |coord| = 0.5, so points that the oracle places within
``1e-3`` of an integer grid coordinate, or within ``1e-4`` of the frustum
faces, are resampled before a case is used for gradient parity.
"""
import torch

from oracle import closed_form as CF


def make_case(cfg, P, N, seed, kind="uniform", translation=False, focal=False,
              scale=True, screened=True):
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        pts = (torch.rand(P, N, 3, generator=g) - 0.5) * 0.9
    elif kind == "clustered":
        pts = torch.clamp(0.15 * torch.randn(P, N, 3, generator=g), -0.5, 0.5)
    elif kind == "recipe":
        pts = torch.rand(P, N, 3, generator=g)
    else:
        raise ValueError(kind)
    quat = torch.randn(P, 4, generator=g)
    case = {"points": pts, "quat": quat,
            "translation": 0.05 * torch.randn(P, 3, generator=g) if translation else None,
            "focal": 1.875 + 0.3 * (torch.rand(P, 1, generator=g) - 0.5) if focal else None,
            "scale": 0.2 + 0.8 * torch.rand(P, 1, generator=g) if scale else None}
    if screened:
        case["points"] = screen(cfg, case, g)
    return case


def near_boundary(cfg, tr_pc, tol_cell=1e-3, tol_face=1e-4):
    vz, v = CF.grid_dims(cfg)
    dims = torch.tensor([vz, v, v], dtype=torch.float64)
    g = (tr_pc + 0.5) * (dims - 1)
    cell = (g - torch.round(g)).abs() < tol_cell
    face = ((tr_pc.abs() - 0.5).abs() < tol_face)
    return (cell | face).any(dim=-1)


def screen(cfg, case, gen, max_iter=50):
    pts = case["points"].clone()
    for _ in range(max_iter):
        tr = CF.pose_transform(cfg, pts, case["quat"], case["translation"], case["focal"])
        bad = near_boundary(cfg, tr)
        n_bad = int(bad.sum())
        if n_bad == 0:
            return pts
        pts[bad] = (torch.rand(n_bad, 3, generator=gen) - 0.5) * 0.9
    raise RuntimeError("screening did not converge")


def loss_weights(P, V, seed=5):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(P, V, V, 1, generator=g), torch.rand(P, V, V, 1, generator=g))
