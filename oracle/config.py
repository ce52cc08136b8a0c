"""Config keys read by the projection path, with the reference's defaults.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Defaults follow /root/reference/dpc/resources/default_config.yaml (lines
27-31 point cloud, 49-56 gauss filter, 72-83 projection/DRC) overlaid with
experiments/chair_unsupervised/config.yaml:7-14 where ``chair_unsupervised``
is asked for.  The reference builds an EasyDict from YAML at import time
(util/config.py:131-155); easydict is not installed here, so this is a plain
attribute dict.
"""


class AttrDict(dict):
    """dict with attribute access (what the reference's EasyDict provides)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return AttrDict(dict.copy(self))


# default_config.yaml values for every key the hot path reads
_DEFAULTS = dict(
    vox_size=64,
    vox_size_z=-1,
    camera_distance=2.0,
    focal_length=1.875,
    pose_quaternion=True,
    pc_separable_gauss_filter=True,
    pc_gauss_kernel_size=11,
    pc_relative_sigma=1.0,
    pc_relative_sigma_end=0.2,
    pc_rgb_stop_points_gradient=False,
    pc_rgb_divide_by_occupancies=False,
    pc_rgb_divide_by_occupancies_epsilon=0.01,
    pc_rgb_clip_after_conv=False,
    ptn_max_projection=False,
    drc_logsum=True,
    drc_logsum_clip_val=0.00001,
    drc_tf_cumulative=True,
    max_depth=10.0,
    max_number_of_steps=600000,
    pc_num_points=8000,
    # candidate-selection loss, default_config.yaml
    pose_predict_num_candidates=1,
    pose_predictor_student=False,
    variable_num_views=False,
    pc_gauss_filter_gt=False,
    proj_weight=1.0,
)

# experiments/chair_unsupervised/config.yaml:7-14
_CHAIR_UNSUPERVISED = dict(
    vox_size=64,
    pc_gauss_kernel_size=21,
    pc_relative_sigma=3.0,
    pc_num_points=8000,
)


def default_cfg(**overrides):
    cfg = AttrDict(_DEFAULTS)
    cfg.update(overrides)
    return cfg


def chair_unsupervised_cfg(**overrides):
    cfg = AttrDict(_DEFAULTS)
    cfg.update(_CHAIR_UNSUPERVISED)
    cfg.update(overrides)
    return cfg
