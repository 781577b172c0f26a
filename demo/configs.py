"""Default configurations of the two demos, with the schema of the reference's yaml files
(demo/pendulum_config.yaml, demo/particle_config.yaml: same sections, keys and values -- a test
compares them with the reference's files where those are available).  `load(path_or_None, default)`
reads any yaml file with that schema instead."""
import copy

PENDULUM = {
    "sim_params": {"episodes": 1, "render": True, "steps": 200, "verbose": False, "warm_up": 0},
    "exp_params": {
        "init_state": [3.0, 0.0], "horizon": 30, "n_particles": 3, "action_samples": 128, "params_samples": 8,
        "alpha": 1, "learning_rate": 2.0, "bandwidth_scaling": 1.0, "ctrl_sigma": 2, "ctrl_dim": 1, "prior_sigma": 2,
        "weighted_prior": False,
        "params_prior_loc": [[0.5, 0.5], [0.5, 1.5], [1.5, 0.5], [1.5, 1.5]], "params_prior_sigma": 0.1,
        "likelihood": "ExponentiatedUtility", "kernel": "rbf",
        "mpf_n_particles": 50, "mpf_steps": 20, "mpf_log_space": False, "mpf_learning_rate": 0.001,
        "mpf_bandwidth": None, "mpf_bandwidth_scaling": 1.0, "mpf_obs_std": 0.1,
    },
    "utf": {"n": 2, "alpha": 0.5},
}

PARTICLE = {
    "sim_params": {"warm_up": 5, "steps": 10, "episodes": 1},
    "exp_params": {
        "horizon": 40, "n_particles": 6, "action_samples": 64, "params_samples": 4, "alpha": 1, "learning_rate": 100,
        "bandwidth_scaling": 1.0, "ctrl_sigma": 5, "ctrl_dim": 2, "likelihood": "ExponentiatedUtility", "sampling": True,
        "kernel": "rbf", "use_svmpc": True, "use_mpf": True, "prior_sigma": 5, "weighted_prior": True,
        "dyn_prior": "Normal", "dyn_prior_arg1": 2, "dyn_prior_arg2": 0.1, "extra_load": 1.0,
        "mpf_n_particles": 50, "mpf_steps": 20, "mpf_log_space": True, "mpf_learning_rate": 0.01, "mpf_bandwidth": 0.5,
        "mpf_bandwidth_scaling": 1.0, "mpf_obs_std": 0.1,
    },
    "env_params": {
        "dt": 0.015, "control_type": "acceleration", "noise_std": [0.1, 0.1], "init_state": [-9.0, -9.0, 0, 0],
        "target_state": [9.0, 9.0, 0, 0], "can_crash": True, "with_obstacle": True, "deterministic": True,
        "cost_params": {"w_qpos": 0.5, "w_qvel": 0.25, "w_ctrl": 0.2, "w_obs": 1.0e6, "w_qpos_T": 1.0e3, "w_qvel_T": 0.1},
        "obst_preset": "grid_4x4", "obst_width": 2.1, "max_speed": 5, "max_accel": 10, "map_cell_size": 0.1,
        "map_size": [22, 22], "map_type": "direct",
    },
}


def load(path, default):
    if path is None:
        return copy.deepcopy(default)
    import yaml

    with open(path) as f:
        return yaml.load(f, yaml.FullLoader)
