"""ctypes binding of lib/libdpc_b200.so (the C ABI in include/dpc_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails the
error is raised to the caller.  Build with ``python __graft_entry__.py`` (or
``make -C pytorch-unsup-pc_b200/csrc``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPC_B200_LIB: load another build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("DPC_B200_LIB") or os.path.join(_HERE, "lib", "libdpc_b200.so")

c_float_p = ctypes.c_void_p   # device/host pointers travel as raw addresses
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_size_t = ctypes.c_size_t

SCATTER_ATOMIC = 0
SCATTER_SORTED = 1
OUT_VOXELS = 1      # dpc_params.outputs flags
OUT_PROBS = 2
MAX_TAPS = 21


class Params(ctypes.Structure):
    """``dpc_params`` (include/dpc_b200.h)."""
    _fields_ = [("P", ctypes.c_int32), ("N", ctypes.c_int32),
                ("Vz", ctypes.c_int32), ("V", ctypes.c_int32),
                ("camera_distance", ctypes.c_double), ("focal_length", ctypes.c_double),
                ("max_depth", ctypes.c_double), ("drc_clip", ctypes.c_double),
                ("drc_logsum", ctypes.c_int32), ("flip_y", ctypes.c_int32),
                ("outputs", ctypes.c_int32)]


_P = ctypes.POINTER(Params)

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "dpc_version": [],
    "dpc_last_error": [],
    "dpc_workspace_bytes": [_P],
    "dpc_cells_bytes": [_P],
    "dpc_project_chunks": [_P],
    "dpc_project_kernels_per_chunk": [_P],
    "dpc_pose_fwd": [_P] + [c_void_p] * 5 + [c_void_p],
    "dpc_pose_bwd": [_P] + [c_void_p] * 9 + [c_void_p, c_size_t, c_void_p],
    "dpc_scatter_fwd": [_P, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p],
    "dpc_scatter_bwd": [_P, c_void_p, c_void_p, c_void_p, c_void_p],
    "dpc_blur3d": [_P, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                   c_void_p],
    "dpc_drc_fwd": [_P] + [c_void_p] * 4 + [c_void_p],
    "dpc_drc_bwd": [_P] + [c_void_p] * 5 + [c_void_p, c_size_t, c_void_p],
    "dpc_depth_from_probs_fwd": [_P, c_void_p, c_void_p, c_void_p],
    "dpc_depth_from_probs_bwd": [_P, c_void_p, c_void_p, c_void_p],
    "dpc_project_fwd": [_P] + [c_void_p] * 5
                       + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int]
                       + [c_void_p] * 8 + [c_void_p, c_size_t, c_void_p],
    "dpc_project_bwd": [_P] + [c_void_p] * 5
                       + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int]
                       + [c_void_p] * 3 + [c_void_p] * 5 + [c_void_p] * 6
                       + [c_void_p, c_size_t, c_void_p],
    "dpc_project_replicated_fwd": [_P, c_int, c_int, c_void_p] + [c_void_p] * 5
                       + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int]
                       + [c_void_p] * 8 + [c_void_p, c_size_t, c_void_p],
    "dpc_project_replicated_bwd": [_P, c_int, c_int, c_void_p] + [c_void_p] * 5
                       + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int]
                       + [c_void_p] * 3 + [c_void_p] * 5 + [c_void_p] * 8
                       + [c_void_p, c_size_t, c_void_p],
    "dpc_point_dropout_indices": [c_int, c_int, c_int, ctypes.c_uint64, c_void_p, c_void_p],
    "dpc_select_points": [c_int] * 5 + [c_void_p] * 3 + [c_void_p],
    "dpc_replica_reduce": [c_int] * 5 + [c_void_p] * 4 + [c_void_p],
    "dpc_feat_scatter_fwd": [_P, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "dpc_feat_scatter_bwd": [_P, c_int] + [c_void_p] * 6 + [c_void_p],
    "dpc_blur3d_clamped": [_P, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                           c_void_p],
    "dpc_colour_fwd": [_P, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float, c_int, c_void_p,
                       c_void_p, c_void_p],
    "dpc_colour_bwd": [_P, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float, c_int, c_void_p,
                       c_void_p, c_void_p, c_void_p],
    "dpc_candidate_loss_fwd": [c_int] * 4 + [c_void_p] * 6 + [c_void_p],
    "dpc_candidate_loss_bwd": [c_int] * 4 + [c_void_p] * 5 + [ctypes.c_float, c_void_p, c_void_p],
    "dpc_render_loss_fwd": [_P, c_int, c_int, c_void_p] + [c_void_p] * 5
                       + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int]
                       + [c_int, c_int, c_void_p, c_void_p, ctypes.c_float]
                       + [c_void_p] * 4 + [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p],
    "dpc_render_loss_bwd": [_P, c_int, c_int, c_void_p] + [c_void_p] * 5
                       + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int]
                       + [c_int, c_int, c_void_p, c_void_p, ctypes.c_float]
                       + [c_void_p] * 8 + [c_void_p] * 4 + [c_void_p] * 5
                       + [c_void_p, c_size_t, c_void_p],
    "dpc_render_loss_slots": [_P, c_int, c_int, c_int],
    "dpc_release": [],
    "dpc_tap_radius": [c_void_p, c_int],
    "dpc_set_tap_truncation": [ctypes.c_double],
    "dpc_set_programmatic_launch": [c_int],
    "dpc_fma_rate_probe": [c_int, c_int, c_void_p, c_void_p, c_void_p],
    "dpc_point_cloud_distance": [c_int, c_int] + [c_void_p] * 5 + [c_void_p, c_size_t, c_void_p],
    "dpc_project_profile": [_P] + [c_void_p] * 5
                           + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int]
                           + [c_void_p] * 6 + [c_void_p] * 2 + [c_void_p] * 6
                           + [c_void_p, c_size_t, c_void_p, c_int, c_void_p],
}
PROFILE_STAGES = ("memset", "pose_scatter", "blur_xy_fwd", "blurz_drc_fwd", "drc_blurz_bwd",
                  "blur_xy_bwd", "gather_pose_bwd")
_RESTYPES = {"dpc_last_error": ctypes.c_char_p, "dpc_workspace_bytes": c_size_t,
             "dpc_cells_bytes": c_size_t}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "dpc_b200: CUDA library %s is missing -- build it with "
            "`python __graft_entry__.py` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().dpc_last_error()
        raise RuntimeError("dpc_b200 %s failed (status %d): %s"
                           % (what, status, msg.decode() if msg else "?"))
