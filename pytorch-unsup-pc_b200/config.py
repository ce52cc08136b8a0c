"""A plain attribute-dict configuration with the reference's defaults for every key the
projection path reads (dpc/resources/default_config.yaml:27-31, 49-56, 61-66, 72-83, 107-114).

The reference builds an EasyDict from YAML (util/config.py:131-155); any object with these
attributes works as ``cfg`` for the functions of this package -- the reference's own EasyDict,
or ``default_cfg(vox_size=64, pc_gauss_kernel_size=21)`` where the reference is not installed.
"""


class Config(dict):
    """dict with attribute access (what the reference's EasyDict provides)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


_DEFAULTS = dict(
    vox_size=64, vox_size_z=-1, camera_distance=2.0, focal_length=1.875, pose_quaternion=True,
    pc_separable_gauss_filter=True, pc_gauss_kernel_size=11, pc_relative_sigma=1.0,
    pc_relative_sigma_end=0.2, pc_rgb_stop_points_gradient=False,
    pc_rgb_divide_by_occupancies=False, pc_rgb_divide_by_occupancies_epsilon=0.01,
    pc_rgb_clip_after_conv=False, ptn_max_projection=False, drc_logsum=True,
    drc_logsum_clip_val=0.00001, drc_tf_cumulative=True, max_depth=10.0,
    max_number_of_steps=600000, pc_num_points=8000, pose_predict_num_candidates=1,
    pose_predictor_student=False, variable_num_views=False, pc_gauss_filter_gt=False,
    proj_weight=1.0,
)


def default_cfg(**overrides):
    cfg = Config(_DEFAULTS)
    cfg.update(overrides)
    return cfg
