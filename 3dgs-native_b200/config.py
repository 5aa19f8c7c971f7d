"""Constants and training configuration of the reference (config.py:8-113), restated without Warp.
Only the values the hot path and the step loop read are kept; keys the reference reads with
``.get(default)`` but never defines (train.py:385-388,644-646,256) carry those defaults here."""

SEED = 42                      # config.py:8
TILE_M = TILE_N = 16           # config.py:21-22
TILE_THREADS = 256             # config.py:23


class GaussianParams:
    num_iterations = 7000
    num_points = 5000
    save_interval = 500
    use_lr_scheduler = True
    lr_scheduler_config = {"lr_pos": 1e-2, "lr_scale": 5e-3, "lr_rot": 5e-3, "lr_sh": 2e-3, "lr_opac": 5e-3,
                           "final_lr_factor": 0.01}
    adam_beta1, adam_beta2, adam_epsilon = 0.9, 0.999, 1e-8
    densification_interval = 100
    opacity_reset_interval = 3000
    densify_grad_threshold = 0.0002
    cull_opacity_threshold = 0.005
    percent_dense = 0.01
    max_allowed_prune_ratio = 1.0
    initial_scale = 0.1
    scale_modifier = 1.0
    sh_degree = 3
    background_color = [0.0, 0.0, 0.0]
    lambda_dssim = 0.0
    # defaults of train.py's config.get(...) calls
    densify_from_iter = 500
    densify_until_iter = 15000
    min_valid_points = 1000
    max_valid_points = 1000000
    camera_extent_factor = 1.0

    @classmethod
    def update(cls, **kwargs):
        for key, value in kwargs.items():
            if not hasattr(cls, key):
                raise ValueError(f"Unknown parameter: {key}")   # config.py:82
            setattr(cls, key, value)

    @classmethod
    def get_config_dict(cls):
        return {k: getattr(cls, k) for k in dir(cls)
                if not k.startswith("_") and not callable(getattr(cls, k))}
