# dpc/util/dpc_b200_binding.py  -- ctypes stub for the reference tree
import ctypes, torch

class Params(ctypes.Structure):            # dpc_params, include/dpc_b200.h
    _fields_ = [("P", ctypes.c_int32), ("N", ctypes.c_int32), ("Vz", ctypes.c_int32),
                ("V", ctypes.c_int32), ("camera_distance", ctypes.c_double),
                ("focal_length", ctypes.c_double), ("max_depth", ctypes.c_double),
                ("drc_clip", ctypes.c_double), ("drc_logsum", ctypes.c_int32),
                ("flip_y", ctypes.c_int32), ("outputs", ctypes.c_int32)]

lib = ctypes.CDLL(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "..", "pytorch-unsup-pc_b200", "lib", "libdpc_b200.so"))
lib.dpc_workspace_bytes.restype = lib.dpc_cells_bytes.restype = ctypes.c_size_t
lib.dpc_last_error.restype = ctypes.c_char_p
vp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())

def project_forward(cfg, points, quat, scale, taps):     # taps: 3 CPU fp32 tensors of K values
    P, N, _ = points.shape; V = cfg.vox_size
    p = Params(P, N, V, V, cfg.camera_distance, cfg.focal_length, cfg.max_depth,
               cfg.drc_logsum_clip_val, 1, 1, 0)     # outputs=0: no voxels / drc_probs tensors
    dev = points.device
    new = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
    tr_pc, grid, bits = new(P, N, 3), new(P, V, V, V), new(P, V, V, V // 32, dt=torch.int32)
    mask, depth = new(P, V, V), new(P, V, V)
    cells = new(lib.dpc_cells_bytes(ctypes.byref(p)), dt=torch.uint8)   # cell records + ray checkpoints
    ws = new(lib.dpc_workspace_bytes(ctypes.byref(p)), dt=torch.uint8)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = lib.dpc_project_fwd(ctypes.byref(p), vp(points), vp(quat), None, None, vp(scale),
                             vp(taps[0]), taps[0].numel(), vp(taps[1]), taps[1].numel(),
                             vp(taps[2]), taps[2].numel(), 0,            # DPC_SCATTER_ATOMIC
                             vp(tr_pc), vp(grid), vp(bits), vp(cells), vp(mask), vp(depth),
                             None, None, vp(ws), ctypes.c_size_t(ws.numel()), stream)
    if rc != 0:
        raise RuntimeError(lib.dpc_last_error().decode())
    return mask, depth, tr_pc, (grid, bits, cells)   # the state dpc_project_bwd needs


if __name__ == "__main__":
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    import pytorch_unsup_pc_b200 as dpc
    cfg = dpc.default_cfg(vox_size=32, pc_gauss_kernel_size=11)
    g = torch.Generator().manual_seed(1)
    pts = ((torch.rand(2, 500, 3, generator=g) - 0.5) * 0.9).cuda()
    quat = torch.randn(2, 4, generator=g).cuda()
    scale = (0.2 + 0.8 * torch.rand(2, generator=g)).cuda()
    kern = dpc.smoothing_kernel(cfg, 1.5)
    taps = [k.reshape(-1).contiguous() for k in kern]
    mask, depth, tr_pc, state = project_forward(cfg, pts, quat, scale, taps)
    dpc.set_outputs(voxels=False, drc_probs=False)
    ref = dpc.pointcloud_project_fast(cfg, pts, quat, None, None, kern, scaling_factor=scale.reshape(-1, 1))
    torch.cuda.synchronize()
    assert torch.equal(mask, ref["proj"].squeeze(-1)) and torch.equal(depth, ref["proj_depth"].squeeze(-1))
    assert torch.equal(tr_pc, ref["tr_pc"])
    print("INTEGRATION.md binding stub: OK (bit-identical to the package's own call)")
