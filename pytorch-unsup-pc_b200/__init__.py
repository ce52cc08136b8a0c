"""dpc_b200: B200-native differentiable point-cloud projection.

Drop-in replacement for the projection path of NiteshBharadwaj/pytorch-unsup-pc
(util/point_cloud_to.py, util/drc.py, util/gauss_kernel.py): same function
names and signatures, hand-written sm_100a CUDA kernels behind a C ABI
(include/dpc_b200.h).  Import as ``pytorch_unsup_pc_b200``.
"""
from . import _lib
from .config import Config, default_cfg
from .gauss_kernel import gauss_kernel_1d, separable_kernels, smoothing_kernel
from .point_cloud import (pc_perspective_transform, pointcloud2voxels3d_fast,
                          pointcloud_project_fast, smoothen_voxels3d, set_outputs,
                          set_deterministic, options, pc_point_dropout, convolve_rgb,
                          pointcloud_project_replicated)
from .drc import (drc_depth_projection, drc_event_probabilities, drc_projection,
                  project_volume_rgb_integral)
from .pipeline import AlternatingGraphs, GraphedSteps, HostPipeline, bind_to_device_numa
from .losses import add_proj_loss, proj_loss_pose_candidates, project_candidates_loss
from .point_cloud_distance import chamfer_distances, point_cloud_distance

__all__ = [
    "pointcloud_project_fast", "pc_perspective_transform", "pointcloud2voxels3d_fast",
    "smoothen_voxels3d", "drc_projection", "drc_depth_projection", "drc_event_probabilities",
    "smoothing_kernel", "gauss_kernel_1d", "separable_kernels",
    "set_outputs", "set_deterministic", "options", "HostPipeline", "GraphedSteps", "AlternatingGraphs", "bind_to_device_numa", "add_proj_loss", "proj_loss_pose_candidates", "project_candidates_loss", "point_cloud_distance", "chamfer_distances",
    "pc_point_dropout", "pointcloud_project_replicated", "convolve_rgb",
    "project_volume_rgb_integral",
    "library_path", "version", "Config", "default_cfg",
]


def library_path():
    return _lib.LIB_PATH


def version():
    return _lib.load().dpc_version()
