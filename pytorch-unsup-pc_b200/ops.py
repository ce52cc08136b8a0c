"""PyTorch custom autograd ops over the C ABI (include/dpc_b200.h).

PyTorch supplies device memory, the current stream and the autograd tape;
every computation is one of the hand-written sm_100a kernels in csrc/.  All
ops require contiguous fp32 CUDA tensors and raise otherwise -- there is no
CPU or eager fallback.
"""
import ctypes

import torch

from . import _lib

_ws_cache = {}
_retired = []        # superseded scratch buffers (see _scratch)


def _ptr(t):
    # a plain int converts to void* as well as a c_void_p object does, and costs no allocation
    return None if t is None else t.data_ptr()


def _stream(device):
    # the raw handle of torch's current stream (torch.cuda.current_stream() builds a
    # Python Stream object every call: ~10 us, which matters at ~200 us per step)
    return torch._C._cuda_getCurrentRawStream(device.index)


class _on_device:
    """``with torch.cuda.device(dev)`` only when ``dev`` is not already current."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if torch._C._cuda_getDevice() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _f32(t, name, shape=None):
    """Validate/normalise one tensor argument (None passes through)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: dpc_b200 has no CPU fallback" % name)
    if t.dtype != torch.float32:
        t = t.float()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t.contiguous()


def make_params(cfg, P, N, flip_y=True, outputs=0):
    """dpc_params from the reference's cfg keys (default_config.yaml:72-83); ``outputs`` =
    the optional outputs (OUT_VOXELS | OUT_PROBS) the whole-path forward materialises."""
    V = int(cfg.vox_size)
    vz = int(getattr(cfg, "vox_size_z", -1))
    Vz = V if vz == -1 else vz
    return _lib.Params(P=int(P), N=int(N), Vz=Vz, V=V,
                       camera_distance=float(cfg.camera_distance),
                       focal_length=float(cfg.focal_length),
                       max_depth=float(getattr(cfg, "max_depth", 10.0)),
                       drc_clip=float(getattr(cfg, "drc_logsum_clip_val", 1e-5)),
                       drc_logsum=1 if getattr(cfg, "drc_logsum", True) else 0,
                       flip_y=1 if flip_y else 0, outputs=int(outputs))


def host_taps(kernel):
    """[k1, k2, k3] conv3d kernels (or None) -> three flat fp32 HOST tensors.

    The taps are kernel *arguments* (they ride in the constant bank), so they
    must be readable by the host: CPU tensors -- what the reference's
    ``smoothing_kernel`` returns -- are used as they are; CUDA tensors cost a
    device-to-host copy and a sync."""
    if kernel is None:
        return None
    if len(kernel) != 3:
        raise ValueError("kernel must be the [X, Y, Z] list smoothing_kernel returns")
    out = []
    for k in kernel:
        if not (isinstance(k, torch.Tensor) and k.dtype == torch.float32 and not k.is_cuda
                and not k.requires_grad and k.is_contiguous()):
            k = torch.as_tensor(k).detach().to(device="cpu", dtype=torch.float32)
        k = k.reshape(-1)
        if k.numel() % 2 == 0 or k.numel() > _lib.MAX_TAPS:
            raise ValueError("Gaussian kernel size %d must be odd and <= %d"
                             % (k.numel(), _lib.MAX_TAPS))
        out.append(k.contiguous())
    return tuple(out)


def _tap_args(taps):
    if taps is None:
        return (None, 0, None, 0, None, 0)
    args = []
    for k in taps:
        args += [k.data_ptr(), int(k.numel())]
    return tuple(args)


_size_cache = {}


def _sizes(params):
    """(workspace bytes, saved cell-record bytes) of a geometry, asked from the library once."""
    key = (params.P, params.N, params.Vz, params.V)
    v = _size_cache.get(key)
    if v is None:
        lib = _lib.load()
        v = (lib.dpc_workspace_bytes(ctypes.byref(params)), lib.dpc_cells_bytes(ctypes.byref(params)))
        _size_cache[key] = v
    return v


def _workspace(params, device):
    """One cached workspace per (device, size class); stream-ordered reuse."""
    return _scratch("ws", _sizes(params)[0], device)


def _scratch(kind, nbytes, device):
    """Cached scratch bytes per (kind, device, stream): contents are dead between
    calls and every use is ordered on that stream, so reuse needs no allocation."""
    key = (kind, device.index, torch._C._cuda_getCurrentRawStream(device.index))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            # a CUDA graph captured over this path (pipeline.GraphedSteps) has the old pointer
            # baked in: a superseded buffer is retired, never freed, so a replay can not write
            # into memory the allocator has handed to another tensor
            _retired.append(buf)
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


class ProjectFn(torch.autograd.Function):
    """pointcloud_project_fast (point_cloud_to.py:191-263) as ONE op.

    forward : pose + cell records, plane scatter + blur XY, blur Z/DRC   (3 launches)
    backward: DRC reverse scan/blur Z adjoint, blur XY adjoint + plane gather,
              pose adjoint with fused final reductions                   (3 launches)
    (the deterministic mode: items + sort in front of the plane kernel, which then sums
    every plane row in a fixed order; ``plane_local=False``: memset + global scatter -- or
    the stand-alone sorted scatter -- forward, grid gather backward.)
    Saved for backward: the inputs, the blurred grid, a 1-bit clamp mask and
    the per-point cell records.
    """

    @staticmethod
    def forward(ctx, points, quat, trans, focal, scale, params, taps, want_voxels, want_probs,
                mode, plane_local=True, rep=None):
        # rep = (replicas, N_src, sel) for the replica-aware entry points (next row f2): points
        # is then the un-replicated [P/replicas, N_src, 3] cloud tensor; None = plain [P,N,3]
        lib = _lib.load()
        dev = points.device
        P, N, Vz, V = params.P, params.N, params.Vz, params.V
        params.outputs = (_lib.OUT_VOXELS if want_voxels else 0) | (_lib.OUT_PROBS if want_probs else 0)
        f32 = dict(dtype=torch.float32, device=dev)
        tr_pc = torch.empty(P, N, 3, **f32)
        mask = torch.empty(P, V, V, **f32)
        depth = torch.empty(P, V, V, **f32)
        voxels = torch.empty(P, Vz, V, V, **f32) if want_voxels else None
        probs = torch.empty(Vz + 1, P, V, V, **f32) if want_probs else None
        # state saved for the backward, ONE allocation: blurred grid | clamp bits | per-point
        # cell records (plane-local scatter/gather path, the default mode)
        use_cells = bool(plane_local)      # both scatter modes save the plane-local state
        n_grid = P * Vz * V * V * 4
        n_bits = P * Vz * V * (V // 32) * 4
        n_cells = _sizes(params)[1] if use_cells else 0
        state = torch.empty(n_grid + n_bits + n_cells, dtype=torch.uint8, device=dev)
        base = state.data_ptr()
        grid_b, bits = base, base + n_grid
        cells = base + n_grid + n_bits if use_cells else None
        ws = _workspace(params, dev)
        tail = (_ptr(points), _ptr(quat), _ptr(trans), _ptr(focal),
                _ptr(scale), *_tap_args(taps), int(mode), _ptr(tr_pc), grid_b, bits,
                cells, _ptr(mask), _ptr(depth), _ptr(voxels), _ptr(probs), _ptr(ws),
                ws.numel(), _stream(dev))
        with _on_device(dev):
            if rep is None:
                st = lib.dpc_project_fwd(ctypes.byref(params), *tail)
            else:
                st = lib.dpc_project_replicated_fwd(ctypes.byref(params), int(rep[0]), int(rep[1]),
                                                    _ptr(rep[2]), *tail)
        _lib.check(st, "project_fwd")
        ctx.save_for_backward(points, quat, trans, focal, scale, state,
                              None if rep is None else rep[2])
        ctx.params, ctx.taps = params, taps
        ctx.rep = None if rep is None else (int(rep[0]), int(rep[1]))
        ctx.state_layout = (n_grid, n_bits, use_cells)
        ctx.set_materialize_grads(False)
        return mask, depth, tr_pc, voxels, probs

    @staticmethod
    def backward(ctx, g_mask, g_depth, g_trpc, g_voxels, g_probs):
        lib = _lib.load()
        points, quat, trans, focal, scale, state, sel = ctx.saved_tensors
        n_grid, n_bits, use_cells = ctx.state_layout
        base = state.data_ptr()
        grid_b, bits = base, base + n_grid
        cells = base + n_grid + n_bits if use_cells else None
        params = ctx.params
        dev = points.device
        P, N, Vz, V = params.P, params.N, params.Vz, params.V
        f32 = dict(dtype=torch.float32, device=dev)
        g_mask = _f32(g_mask, "g_mask", (P, V, V))
        g_depth = _f32(g_depth, "g_depth", (P, V, V))
        g_trpc = _f32(g_trpc, "g_tr_pc", (P, N, 3))
        g_voxels = _f32(g_voxels, "g_voxels", (P, Vz, V, V))
        g_probs = _f32(g_probs, "g_probs", (Vz + 1, P, V, V))
        g_grid = _scratch("g_grid", n_grid, dev)      # dead after this call
        g_points = torch.empty(P, N, 3, **f32) if ctx.rep is None else None
        g_quat = torch.empty(P, 4, **f32)
        g_trans = torch.empty(P, 3, **f32) if trans is not None else None
        g_focal = torch.empty(P, **f32) if focal is not None else None
        g_scale = torch.empty(P, **f32) if scale is not None else None
        ws = _workspace(params, dev)
        head = (_ptr(points), _ptr(quat), _ptr(trans), _ptr(focal),
                _ptr(scale), *_tap_args(ctx.taps), grid_b, bits, cells,
                _ptr(g_mask), _ptr(g_depth), _ptr(g_probs), _ptr(g_voxels), _ptr(g_trpc), _ptr(g_grid))
        tail = (_ptr(g_quat), _ptr(g_trans), _ptr(g_focal), _ptr(g_scale),
                _ptr(ws), ws.numel(), _stream(dev))
        with _on_device(dev):
            if ctx.rep is None:
                st = lib.dpc_project_bwd(ctypes.byref(params), *head, _ptr(g_points), *tail)
            else:
                # per-replica point gradients are scratch; the gradient of the CLOUD tensor is
                # their sum over the replicas routed through the dropout selection
                R, N_src = ctx.rep
                g_rep = _scratch("g_rep", P * N * 12, dev)
                inv = _scratch("inv", P * N_src * 4, dev) if sel is not None else None
                g_points = torch.empty(P // R, N_src, 3, **f32)
                st = lib.dpc_project_replicated_bwd(
                    ctypes.byref(params), R, N_src, _ptr(sel), *head, _ptr(g_rep), _ptr(inv),
                    _ptr(g_points), *tail)
        _lib.check(st, "project_bwd")
        return (g_points, g_quat, g_trans, g_focal, g_scale) + (None,) * 7


def project(points, quat, trans, focal, scale, params, taps, want_voxels, want_probs, mode,
            plane_local=True, rep=None):
    """The whole-path op: through the C++ autograd binding when lib/dpc_b200_torch.so is built
    (no Python in either pass), else through ``ProjectFn`` (ctypes).  Same C ABI, same kernels,
    same results.  Returns (mask, depth, tr_pc, voxels or None, probs or None)."""
    from . import _torch_binding
    ext = _torch_binding.load()
    if ext is None:
        return ProjectFn.apply(points, quat, trans, focal, scale, params, taps, want_voxels,
                               want_probs, mode, plane_local, rep)
    geom = [params.P, params.N, params.Vz, params.V, params.camera_distance, params.focal_length,
            params.max_depth, params.drc_clip, params.drc_logsum, params.flip_y]
    kx, ky, kz = taps if taps is not None else (None, None, None)
    replicas, n_src, sel = (0, 0, None) if rep is None else (int(rep[0]), int(rep[1]), rep[2])
    out = ext.project(points, quat, trans, focal, scale, geom, kx, ky, kz, bool(want_voxels),
                      bool(want_probs), int(mode), bool(plane_local), replicas, n_src, sel)
    mask, depth, tr_pc = out[:3]
    voxels = out[3] if want_voxels else None
    probs = out[3 + bool(want_voxels)] if want_probs else None
    return mask, depth, tr_pc, voxels, probs


class RenderLossFn(torch.autograd.Function):
    """Replica-aware projection + candidate-selection loss as ONE op (model_pc_to.py:302-331 +
    :339-385, 410-440): clouds, poses and ground-truth masks in, the loss out; the backward runs
    over the winning candidate of every view only (the others' gradients are exactly zero) with
    dL/dmask built inside the ray kernel.  Returns (loss [], min_idx [BV], all_loss [BV,C],
    mask [P,V,V]); only ``loss`` is differentiable."""

    @staticmethod
    def forward(ctx, points, quat, trans, focal, scale, masks, weights, params, taps, C,
                weight_scale, mode, plane_local, replicas, N_src, sel):
        lib = _lib.load()
        dev = points.device
        P, N, Vz, V = params.P, params.N, params.Vz, params.V
        BV, G = masks.shape[0], masks.shape[-1]
        params.outputs = 0
        f32 = dict(dtype=torch.float32, device=dev)
        mask = torch.empty(P, V, V, **f32)
        all_loss = torch.empty(BV, C, **f32)
        min_idx = torch.empty(BV, dtype=torch.int64, device=dev)
        # view_loss [BV] | kcoef [BV] | loss [1] (fp32) and winners [BV] (int32), one allocation
        small = torch.empty(3 * BV + 1, **f32)
        view_loss, kcoef, loss = small[:BV], small[BV:2 * BV], small[2 * BV:2 * BV + 1]
        winners = small[2 * BV + 1:].view(torch.int32)
        use_cells = bool(plane_local)      # both scatter modes save the plane-local state
        n_grid = P * Vz * V * V * 4
        n_bits = P * Vz * V * (V // 32) * 4
        n_cells = _sizes(params)[1] if use_cells else 0
        state = torch.empty(n_grid + n_bits + n_cells, dtype=torch.uint8, device=dev)
        base = state.data_ptr()
        grid_b, bits = base, base + n_grid
        cells = base + n_grid + n_bits if use_cells else None
        ws = _workspace(params, dev)
        with _on_device(dev):
            st = lib.dpc_render_loss_fwd(
                ctypes.byref(params), int(replicas), int(N_src), _ptr(sel), _ptr(points), _ptr(quat),
                _ptr(trans), _ptr(focal), _ptr(scale), *_tap_args(taps), int(mode), int(C), int(G),
                _ptr(masks), _ptr(weights), ctypes.c_float(weight_scale), grid_b, bits, cells,
                _ptr(mask), _ptr(all_loss), _ptr(min_idx), _ptr(view_loss), _ptr(loss),
                _ptr(winners), _ptr(kcoef), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(st, "render_loss_fwd")
        ctx.save_for_backward(points, quat, trans, focal, scale, masks, weights, sel, state, mask,
                              min_idx, small)
        ctx.params, ctx.taps = params, taps
        ctx.meta = (int(C), int(G), float(weight_scale), int(mode), int(replicas), int(N_src),
                    n_grid, n_bits, use_cells)
        ctx.mark_non_differentiable(min_idx, all_loss, mask)
        return loss.reshape(()), min_idx, all_loss, mask

    @staticmethod
    def backward(ctx, g_loss, _g_idx, _g_all, _g_mask):
        lib = _lib.load()
        (points, quat, trans, focal, scale, masks, weights, sel, state, mask, min_idx,
         small) = ctx.saved_tensors
        C, G, weight_scale, mode, replicas, N_src, n_grid, n_bits, use_cells = ctx.meta
        params = ctx.params
        dev = points.device
        P, N, Vz, V = params.P, params.N, params.Vz, params.V
        BV = P // C
        kcoef, winners = small[BV:2 * BV], small[2 * BV + 1:].view(torch.int32)
        base = state.data_ptr()
        grid_b, bits = base, base + n_grid
        cells = base + n_grid + n_bits if use_cells else None
        f32 = dict(dtype=torch.float32, device=dev)
        upstream = _f32(g_loss, "grad_output").reshape(1)
        slots = lib.dpc_render_loss_slots(ctypes.byref(params), C, 1 if use_cells else 0, mode)
        g_grid = _scratch("g_grid", slots * Vz * V * V * 4, dev)
        g_rep = _scratch("g_rep", slots * N * 12, dev)
        inv = _scratch("inv", slots * N_src * 4, dev) if sel is not None else None
        g_mask = _scratch("g_mask", P * V * V * 4, dev) if slots == P else None
        g_points = torch.empty(points.shape, **f32)
        g_quat = torch.empty(P, 4, **f32)
        g_trans = torch.empty(P, 3, **f32) if trans is not None else None
        g_focal = torch.empty(P, **f32) if focal is not None else None
        g_scale = torch.empty(P, **f32) if scale is not None else None
        ws = _workspace(params, dev)
        with _on_device(dev):
            st = lib.dpc_render_loss_bwd(
                ctypes.byref(params), replicas, N_src, _ptr(sel), _ptr(points), _ptr(quat),
                _ptr(trans), _ptr(focal), _ptr(scale), *_tap_args(ctx.taps), mode, C, G,
                _ptr(masks), _ptr(weights), ctypes.c_float(weight_scale), grid_b, bits, cells,
                _ptr(mask), _ptr(min_idx), _ptr(winners), _ptr(kcoef), _ptr(upstream),
                _ptr(g_grid), _ptr(g_rep), _ptr(inv), _ptr(g_mask), _ptr(g_points), _ptr(g_quat),
                _ptr(g_trans), _ptr(g_focal), _ptr(g_scale), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(st, "render_loss_bwd")
        return (g_points, g_quat, g_trans, g_focal, g_scale) + (None,) * 11


class PoseFn(torch.autograd.Function):
    """pc_perspective_transform (point_cloud_to.py:118-178, quaternion branch)."""

    @staticmethod
    def forward(ctx, points, quat, trans, focal, params):
        lib = _lib.load()
        dev = points.device
        tr_pc = torch.empty(params.P, params.N, 3, dtype=torch.float32, device=dev)
        with _on_device(dev):
            st = lib.dpc_pose_fwd(ctypes.byref(params), _ptr(points), _ptr(quat), _ptr(trans),
                                  _ptr(focal), _ptr(tr_pc), _stream(dev))
        _lib.check(st, "pose_fwd")
        ctx.save_for_backward(points, quat, trans, focal)
        ctx.params = params
        return tr_pc

    @staticmethod
    def backward(ctx, g_trpc):
        lib = _lib.load()
        points, quat, trans, focal = ctx.saved_tensors
        params = ctx.params
        dev = points.device
        f32 = dict(dtype=torch.float32, device=dev)
        g_trpc = _f32(g_trpc, "g_tr_pc", (params.P, params.N, 3))
        g_points = torch.empty(params.P, params.N, 3, **f32)
        g_quat = torch.empty(params.P, 4, **f32)
        g_trans = torch.empty(params.P, 3, **f32) if trans is not None else None
        g_focal = torch.empty(params.P, **f32) if focal is not None else None
        ws = _workspace(params, dev)
        with _on_device(dev):
            st = lib.dpc_pose_bwd(ctypes.byref(params), _ptr(points), _ptr(quat), _ptr(trans),
                                  _ptr(focal), _ptr(g_trpc), _ptr(g_points), _ptr(g_quat),
                                  _ptr(g_trans), _ptr(g_focal), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(st, "pose_bwd")
        return g_points, g_quat, g_trans, g_focal, None


class ScatterFn(torch.autograd.Function):
    """pointcloud2voxels3d_fast (point_cloud_to.py:10-87): tr_pc -> raw grid."""

    @staticmethod
    def forward(ctx, tr_pc, params, mode):
        lib = _lib.load()
        dev = tr_pc.device
        grid = torch.empty(params.P, params.Vz, params.V, params.V, dtype=torch.float32, device=dev)
        ws = _workspace(params, dev)
        with _on_device(dev):
            st = lib.dpc_scatter_fwd(ctypes.byref(params), _ptr(tr_pc), _ptr(grid), int(mode),
                                     _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(st, "scatter_fwd")
        ctx.save_for_backward(tr_pc)
        ctx.params = params
        return grid

    @staticmethod
    def backward(ctx, g_grid):
        lib = _lib.load()
        (tr_pc,) = ctx.saved_tensors
        params = ctx.params
        dev = tr_pc.device
        g_grid = _f32(g_grid, "g_grid", (params.P, params.Vz, params.V, params.V))
        g_trpc = torch.empty_like(tr_pc)
        with _on_device(dev):
            st = lib.dpc_scatter_bwd(ctypes.byref(params), _ptr(tr_pc), _ptr(g_grid), _ptr(g_trpc),
                                     _stream(dev))
        _lib.check(st, "scatter_bwd")
        return g_trpc, None, None


def _blur3d(x, params, taps):
    lib = _lib.load()
    dev = x.device
    out = torch.empty_like(x)
    with _on_device(dev):
        st = lib.dpc_blur3d(ctypes.byref(params), _ptr(x), _ptr(out), *_tap_args(taps),
                            _stream(dev))
    _lib.check(st, "blur3d")
    return out


class BlurFn(torch.autograd.Function):
    """smoothen_voxels3d (point_cloud_to.py:90-103).  Symmetric taps with zero
    'same' padding are self-adjoint, so backward is the same three passes."""

    @staticmethod
    def forward(ctx, vox, params, taps):
        ctx.params, ctx.taps = params, taps
        return _blur3d(vox, params, taps)

    @staticmethod
    def backward(ctx, g):
        params = ctx.params
        g = _f32(g, "g_voxels", (params.P, params.Vz, params.V, params.V))
        # adjoint of a cross-correlation = correlation with the reversed taps
        taps = tuple(torch.flip(k, [0]).contiguous() for k in ctx.taps)
        return _blur3d(g, params, taps), None, None


class DrcFn(torch.autograd.Function):
    """drc_projection (drc.py:114-129) + depth (drc.py:152-160): voxels ->
    (mask, probs, depth); no blur, no scaling, flips as asked in params."""

    @staticmethod
    def forward(ctx, vox, params):
        lib = _lib.load()
        dev = vox.device
        P, Vz, V = params.P, params.Vz, params.V
        f32 = dict(dtype=torch.float32, device=dev)
        mask = torch.empty(P, V, V, **f32)
        depth = torch.empty(P, V, V, **f32)
        probs = torch.empty(Vz + 1, P, V, V, **f32)
        with _on_device(dev):
            st = lib.dpc_drc_fwd(ctypes.byref(params), _ptr(vox), _ptr(mask), _ptr(depth),
                                 _ptr(probs), _stream(dev))
        _lib.check(st, "drc_fwd")
        ctx.save_for_backward(vox)
        ctx.params = params
        ctx.set_materialize_grads(False)
        return mask, probs, depth

    @staticmethod
    def backward(ctx, g_mask, g_probs, g_depth):
        lib = _lib.load()
        (vox,) = ctx.saved_tensors
        params = ctx.params
        dev = vox.device
        P, Vz, V = params.P, params.Vz, params.V
        g_mask = _f32(g_mask, "g_mask", (P, V, V))
        g_depth = _f32(g_depth, "g_depth", (P, V, V))
        g_probs = _f32(g_probs, "g_probs", (Vz + 1, P, V, V))
        g_vox = torch.empty_like(vox)
        ws = _workspace(params, dev)
        with _on_device(dev):
            st = lib.dpc_drc_bwd(ctypes.byref(params), _ptr(vox), _ptr(g_mask), _ptr(g_depth),
                                 _ptr(g_probs), _ptr(g_vox), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(st, "drc_bwd")
        return g_vox, None


class DepthFromProbsFn(torch.autograd.Function):
    """drc_depth_projection on stored probabilities (drc.py:152-160)."""

    @staticmethod
    def forward(ctx, probs, params):
        lib = _lib.load()
        dev = probs.device
        depth = torch.empty(params.P, params.V, params.V, dtype=torch.float32, device=dev)
        with _on_device(dev):
            st = lib.dpc_depth_from_probs_fwd(ctypes.byref(params), _ptr(probs), _ptr(depth),
                                              _stream(dev))
        _lib.check(st, "depth_from_probs_fwd")
        ctx.params = params
        return depth

    @staticmethod
    def backward(ctx, g_depth):
        lib = _lib.load()
        params = ctx.params
        dev = g_depth.device
        g_depth = _f32(g_depth, "g_depth", (params.P, params.V, params.V))
        g_probs = torch.empty(params.Vz + 1, params.P, params.V, params.V, dtype=torch.float32,
                              device=dev)
        with _on_device(dev):
            st = lib.dpc_depth_from_probs_bwd(ctypes.byref(params), _ptr(g_depth), _ptr(g_probs),
                                              _stream(dev))
        _lib.check(st, "depth_from_probs_bwd")
        return g_probs, None


def dropout_indices(P, N_src, M, seed, device):
    """sel [P,M] int32: a uniformly random M-subset of range(N_src) per projection, ascending,
    a pure function of (seed, projection index) -- the device replacement of the reference's
    numpy sampler (point_cloud_to.py:275-283)."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("dropout_indices needs a CUDA device: dpc_b200 has no CPU fallback")
    sel = torch.empty(P, M, dtype=torch.int32, device=device)
    with _on_device(device):
        st = lib.dpc_point_dropout_indices(int(P), int(N_src), int(M),
                                           ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(sel),
                                           _stream(device))
    _lib.check(st, "point_dropout_indices")
    return sel


class SelectPointsFn(torch.autograd.Function):
    """select_3d (point_cloud_to.py:266-267) on a cloud tensor shared by `replicas`
    projections: out [P,M,C] = data [P/replicas, N_src, C] at sel [P,M]; backward sums over the
    replicas in replica order (deterministic)."""

    @staticmethod
    def forward(ctx, data, sel, replicas):
        lib = _lib.load()
        dev = data.device
        B, N_src, C = data.shape
        P, M = sel.shape
        out = torch.empty(P, M, C, dtype=torch.float32, device=dev)
        with _on_device(dev):
            st = lib.dpc_select_points(P, int(replicas), N_src, M, C, _ptr(data), _ptr(sel),
                                       _ptr(out), _stream(dev))
        _lib.check(st, "select_points")
        ctx.save_for_backward(sel)
        ctx.dims = (P, int(replicas), N_src, M, C)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        (sel,) = ctx.saved_tensors
        P, R, N_src, M, C = ctx.dims
        dev = sel.device
        g_out = _f32(g_out, "g_out", (P, M, C))
        inv = _scratch("inv", P * N_src * 4, dev)
        g = torch.empty(P // R, N_src, C, dtype=torch.float32, device=dev)
        with _on_device(dev):
            st = lib.dpc_replica_reduce(P, R, N_src, M, C, _ptr(g_out), _ptr(sel), _ptr(inv),
                                        _ptr(g), _stream(dev))
        _lib.check(st, "replica_reduce")
        return g, None, None


class FeatGridFn(torch.autograd.Function):
    """Feature scatter (+ clip + per-channel blur): tr_pc [P,N,3], feat [P,N,C] -> planar
    feature grid [P,C,Vz,V,V]  (TF original point_cloud.py:99-129, 244-249, 148-154).
    ``clip``: clamp(raw, 0, 1) before the blur; its gate is applied inside the gather of the
    backward.  ``taps`` None: no clip, no blur."""

    @staticmethod
    def forward(ctx, tr_pc, feat, params, taps, clip):
        lib = _lib.load()
        dev = tr_pc.device
        P, N, Vz, V = params.P, params.N, params.Vz, params.V
        C = feat.shape[-1]
        raw = torch.empty(P, C, Vz, V, V, dtype=torch.float32, device=dev)
        with _on_device(dev):
            st = lib.dpc_feat_scatter_fwd(ctypes.byref(params), C, _ptr(tr_pc), _ptr(feat), _ptr(raw),
                                          _stream(dev))
        _lib.check(st, "feat_scatter_fwd")
        out = raw
        pc = None
        if taps is not None:
            pc = _lib.Params.from_buffer_copy(params)
            pc.P = P * C
            out = torch.empty_like(raw)
            fn = lib.dpc_blur3d_clamped if clip else lib.dpc_blur3d
            with _on_device(dev):
                st = fn(ctypes.byref(pc), _ptr(raw), _ptr(out), *_tap_args(taps), _stream(dev))
            _lib.check(st, "blur3d(features)")
        gated = bool(clip and taps is not None)
        ctx.save_for_backward(tr_pc, feat, raw if gated else None)
        ctx.params, ctx.pc, ctx.taps = params, pc, taps
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        tr_pc, feat, raw = ctx.saved_tensors
        params = ctx.params
        dev = tr_pc.device
        C = feat.shape[-1]
        g = _f32(g_out, "g_fgrid", (params.P, C, params.Vz, params.V, params.V))
        if ctx.taps is not None:
            rev = tuple(torch.flip(k, [0]).contiguous() for k in ctx.taps)
            gb = torch.empty_like(g)
            with _on_device(dev):
                st = lib.dpc_blur3d(ctypes.byref(ctx.pc), _ptr(g), _ptr(gb), *_tap_args(rev),
                                    _stream(dev))
            _lib.check(st, "blur3d(feature adjoint)")
            g = gb
        g_feat = torch.empty_like(feat)
        g_trpc = torch.empty_like(tr_pc) if ctx.needs_input_grad[0] else None
        with _on_device(dev):
            st = lib.dpc_feat_scatter_bwd(ctypes.byref(params), C, _ptr(tr_pc), _ptr(feat), _ptr(g),
                                          _ptr(raw), _ptr(g_feat), _ptr(g_trpc), _stream(dev))
        _lib.check(st, "feat_scatter_bwd")
        return g_trpc, g_feat, None, None, None


class ColourFn(torch.autograd.Function):
    """Colour integral along the rays (drc.py:132-142 project_volume_rgb_integral) with the
    optional division by the blurred occupancy and clip-after-blur of point_cloud.py:256-262
    applied on the fly: probs [Vz+1,P,V,V], fgrid [P,C,Vz,V,V] (+ div [P,Vz,V,V]) ->
    proj_rgb [P,V,V,C] and, on request, voxels_rgb [P,Vz,V,V,C] (not differentiable)."""

    @staticmethod
    def forward(ctx, probs, fgrid, div, eps, clip_after, params, want_voxels):
        lib = _lib.load()
        dev = probs.device
        P, Vz, V = params.P, params.Vz, params.V
        C = fgrid.shape[1]
        proj = torch.empty(P, V, V, C, dtype=torch.float32, device=dev)
        vox = torch.empty(P, Vz, V, V, C, dtype=torch.float32, device=dev) if want_voxels else None
        with _on_device(dev):
            st = lib.dpc_colour_fwd(ctypes.byref(params), C, _ptr(probs), _ptr(fgrid), _ptr(div),
                                    ctypes.c_float(eps), int(bool(clip_after)), _ptr(proj), _ptr(vox),
                                    _stream(dev))
        _lib.check(st, "colour_fwd")
        ctx.save_for_backward(probs, fgrid, div)
        ctx.params, ctx.eps, ctx.clip_after = params, float(eps), bool(clip_after)
        if vox is not None:
            ctx.mark_non_differentiable(vox)
        return proj, vox

    @staticmethod
    def backward(ctx, g_proj, _g_vox):
        lib = _lib.load()
        probs, fgrid, div = ctx.saved_tensors
        params = ctx.params
        dev = probs.device
        C = fgrid.shape[1]
        g_proj = _f32(g_proj, "g_proj_rgb", (params.P, params.V, params.V, C))
        g_probs = torch.empty_like(probs)
        g_fgrid = torch.empty_like(fgrid)
        with _on_device(dev):
            st = lib.dpc_colour_bwd(ctypes.byref(params), C, _ptr(probs), _ptr(fgrid), _ptr(div),
                                    ctypes.c_float(ctx.eps), int(ctx.clip_after), _ptr(g_proj),
                                    _ptr(g_probs), _ptr(g_fgrid), _stream(dev))
        _lib.check(st, "colour_bwd")
        return g_probs, g_fgrid, None, None, None, None, None
