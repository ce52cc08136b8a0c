"""Torch-on-CPU restatement of the reference projection path (the oracle).

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

Every function cites the reference file:line (relative to /root/reference/dpc)
it restates.  The arithmetic follows the reference's *actual* dtype flow
(SURVEY.md Appendix A): the quaternion normalise and the first Hamilton
product run in the input dtype (fp32); ``quaternion_conjugate`` multiplies by
a float64 numpy array (util/quaternion.py:91-93), so the second product and
everything downstream is fp64.  The same torch op classes are used for the
heavy stages (``index_put_(accumulate=True)``, fp64 ``conv3d``, ``cumsum``),
so timing this module on host cores is a fair stand-in for the reference's
own CPU implementation (bench.py ``cpu_baseline.kind == "port"``).

Gradients: torch autograd over this restatement -- the same mechanism the
reference relies on.

Pinned against the reference executed live: tests/test_oracle_pinning.py, and
against the committed fixtures tests/golden/*.npz made by
tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# a5: Gaussian taps (util/gauss_kernel.py:5-11, 27-32, 35-55)
# --------------------------------------------------------------------------
def gauss_taps(size, sigma):
    """1-D normalised Gaussian, support -size//2+1 .. size//2 (gauss_kernel.py:5-11).

    fp32, as the reference computes it (torch.arange with float bounds)."""
    lo = -size // 2 + 1.0
    hi = size // 2 + 1
    xs = torch.arange(lo, hi)
    k = torch.exp(-xs ** 2 / (2.0 * sigma ** 2))
    return k / k.sum()


def smoothing_taps(cfg, sigma):
    """The three 5-D conv3d kernels [X, Y, Z] (gauss_kernel.py:35-55).

    Anisotropic Z taps when ``vox_size_z != -1`` (:38-51)."""
    size = cfg.pc_gauss_kernel_size
    k = gauss_taps(size, sigma)
    if cfg.vox_size_z != -1:
        ratio = cfg.vox_size_z / cfg.vox_size
        size_z = int((size * ratio) // 1)
        if size_z % 2 == 0:
            size_z += 1
        kz = gauss_taps(size_z, sigma * ratio)
    else:
        kz, size_z = k, size
    return [k.reshape(1, 1, 1, 1, size), k.reshape(1, 1, 1, size, 1),
            kz.reshape(1, 1, size_z, 1, 1)]


# --------------------------------------------------------------------------
# a1 + a2: pose transform (util/quaternion.py:69-132, util/point_cloud_to.py:118-178)
# --------------------------------------------------------------------------
def pose_transform(cfg, points, quat, translation=None, focal_length=None):
    """points [P,N,3] f32, quat [P,4] f32 (unnormalised) -> tr_pc [P,N,3] f64 (z,y,x)."""
    # quaternion.py:119-121 -- normalise in the input dtype
    qn = quat / quat.norm(p=2, dim=-1).reshape(-1, 1)
    qw, qx, qy, qz = [c.unsqueeze(1) for c in qn.unbind(-1)]        # [P,1]
    p0, p1, p2 = points.unbind(-1)                                   # [P,N]
    # quaternion.py:80-85 with b=(0,p): first product, input dtype.  The
    # dropped ``*0`` terms are exact zeros, so rounding is unchanged.
    aw = -(qx * p0) - qy * p1 - qz * p2
    ax = qw * p0 + qy * p2 - qz * p1
    ay = qw * p1 + qz * p0 - qx * p2
    az = qw * p2 + qx * p1 - qy * p0
    # quaternion.py:91-93 -- conjugate via a float64 array => fp64 from here
    bw, bx, by, bz = qw.double(), -qx.double(), -qy.double(), -qz.double()
    aw, ax, ay, az = aw.double(), ax.double(), ay.double(), az.double()
    r0 = aw * bx + ax * bw + ay * bz - az * by
    r1 = aw * by + ay * bw + az * bx - ax * bz
    r2 = aw * bz + az * bw + ax * by - ay * bx
    # point_cloud_to.py:137-139
    if translation is not None:
        t = translation.unsqueeze(1)
        r0 = r0 + t[..., 0]
        r1 = r1 + t[..., 1]
        r2 = r2 + t[..., 2]
    # point_cloud_to.py:130-133, 145-148, 169-175
    f = cfg.focal_length if focal_length is None else focal_length.reshape(-1, 1)
    zc = r0 + cfg.camera_distance
    ys = (r1 * f) / zc
    xs = (r2 * f) / zc
    zs = zc - cfg.camera_distance
    if translation is not None:
        zs = zs - translation.unsqueeze(1)[..., 0]
    return torch.stack([zs, ys, xs], dim=2)


# --------------------------------------------------------------------------
# a3: trilinear scatter (util/point_cloud_to.py:10-87)
# --------------------------------------------------------------------------
def grid_dims(cfg):
    v = cfg.vox_size
    vz = cfg.vox_size_z if cfg.vox_size_z != -1 else v
    return vz, v


def scatter_trilinear(cfg, tr_pc, drop_oob=False):
    """tr_pc [P,N,3] -> raw occupancy grid [P,Vz,V,V] fp64.

    Points with any coordinate outside [-0.5, 0.5] are dropped (:25-27, 42-51).
    A coordinate of exactly +0.5 makes the reference index cell V (IndexError,
    SURVEY.md Appendix A); ``drop_oob=True`` drops those zero-weight corners
    instead, which is what the CUDA path does."""
    vz, v = grid_dims(cfg)
    P, N, _ = tr_pc.shape
    dims = torch.tensor([vz, v, v], dtype=torch.int64)
    inside = ((tr_pc >= -0.5) & (tr_pc <= 0.5)).all(dim=-1)          # [P,N]
    g = (tr_pc + 0.5) * (dims - 1)                                    # :30
    base = torch.floor(g)
    frac = g - base
    cell = base.long()
    b_idx = torch.arange(P).unsqueeze(1).expand(P, N)[inside]
    cell = cell[inside]                                               # [M,3]
    frac = frac[inside]
    grid = torch.zeros(P, vz, v, v, dtype=torch.float64)
    for dz in (0, 1):                                                 # :79-83
        for dy in (0, 1):
            for dx in (0, 1):
                w = ((frac[:, 0] if dz else 1.0 - frac[:, 0])
                     * (frac[:, 1] if dy else 1.0 - frac[:, 1])
                     * (frac[:, 2] if dx else 1.0 - frac[:, 2]))
                iz, iy, ix = cell[:, 0] + dz, cell[:, 1] + dy, cell[:, 2] + dx
                if drop_oob:
                    ok = (iz < vz) & (iy < v) & (ix < v)
                    grid = grid.index_put((b_idx[ok], iz[ok], iy[ok], ix[ok]), w[ok],
                                          accumulate=True)
                else:
                    grid = grid.index_put((b_idx, iz, iy, ix), w, accumulate=True)
    return grid


# --------------------------------------------------------------------------
# a6: separable blur (util/point_cloud_to.py:90-103)
# --------------------------------------------------------------------------
def blur3d(vox, kernels):
    """vox [P,1,Vz,V,V]; three zero-padded 'same' cross-correlations X, Y, Z in fp64."""
    for k in kernels:
        pad = tuple(int(s) // 2 for s in k.shape[2:])
        vox = F.conv3d(vox, k.double(), stride=1, padding=pad)
    return vox


# --------------------------------------------------------------------------
# a8-a10: DRC (util/drc.py:48-106, 114-129, 145-160)
# --------------------------------------------------------------------------
def drc_probabilities(vox, cfg):
    """vox [P,Z,Y,X,1] -> ray-termination probabilities [Z+1,P,Y,X,1].

    Log-sum form (:56-66, 80, 96-104): the 'unity' padding is ``clip_val``,
    not 0 (:59-60), so p_0 and p_Z carry a factor exp(clip_val).  Product form
    (``drc_logsum: false``, :69-71, 82): no clip, true unity."""
    x = vox.permute(1, 0, 2, 3, 4)
    c = cfg.drc_logsum_clip_val
    if cfg.drc_logsum:
        x = torch.clamp(x, c, 1.0 - c)
        log_occ, log_free = torch.log(x), torch.log(1.0 - x)
        run = log_free.cumsum(0)
        pad = torch.full_like(x[:1], c)
        return torch.exp(torch.cat([pad, run], 0) + torch.cat([log_occ, pad], 0))
    run = (1.0 - x).cumprod(0)
    one = torch.ones_like(x[:1])
    return torch.cat([one, run], 0) * torch.cat([x, one], 0)


def drc_mask(probs):
    """Silhouette = sum of all termination events but the last ('escaped') (:121-127)."""
    return probs[:-1].sum(0)


def drc_depth(probs, cfg):
    """Expected depth (:145-160): psi_k = k/Z - 0.5 + camera_distance, psi_Z = max_depth."""
    z = probs.shape[0] - 1
    psi = torch.arange(0, z, 1, dtype=torch.float64) / z - 0.5 + cfg.camera_distance
    psi = torch.cat([psi, torch.tensor([cfg.max_depth], dtype=torch.float64)])
    return (probs * psi.reshape(-1, 1, 1, 1, 1)).sum(0)


# --------------------------------------------------------------------------
# a12: the whole path (util/point_cloud_to.py:191-263, CUDA-branch order)
# --------------------------------------------------------------------------
def project(cfg, points, quat, translation=None, kernels=None, scaling_factor=None,
            focal_length=None, drop_oob=False):
    """Returns the reference's output dict (fp64): proj [P,V,V,1], voxels
    [P,Vz,V,V,1], tr_pc [P,N,3], drc_probs [Vz+1,P,V,V,1], proj_depth [P,V,V,1].

    ``kernels=None`` = no blur (TF original point_cloud.py:237-243)."""
    tr_pc = pose_transform(cfg, points, quat, translation, focal_length)
    raw = scatter_trilinear(cfg, tr_pc, drop_oob=drop_oob)
    vox = torch.clamp(raw.unsqueeze(1), 0.0, 1.0)                     # :198-201
    if kernels is not None:
        vox = blur3d(vox, kernels)                                    # :208
    vox = vox.squeeze(1).unsqueeze(-1)                                # :209
    if scaling_factor is not None:                                    # :218-222
        vox = torch.clamp(vox * scaling_factor.reshape(-1, 1, 1, 1, 1), 0.0, 1.0)
    probs = drc_probabilities(vox, cfg)                               # :238
    mask = drc_mask(probs)
    probs = torch.flip(probs, [2])                                    # :239
    depth = drc_depth(probs, cfg)                                     # :240
    mask = torch.flip(mask, [1])                                      # :242
    return {"proj": mask, "voxels": vox, "tr_pc": tr_pc, "voxels_rgb": None,
            "proj_rgb": None, "drc_probs": probs, "proj_depth": depth,
            "voxels_raw": raw}
