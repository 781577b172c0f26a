#!/usr/bin/env python
"""Inverted-pendulum swing-up with the four controllers the reference's demo compares
(demo/pendulum_example.py:161-263): DuSt-MPC (SVMPC + parameter filter), SVMPC with the belief's mean
parameters, an MPPI baseline that knows the true parameters, and DISCO (one policy, sigma-point
rollouts).  No gym, no plotting: the plant is `PendulumModel.step` with the episode's true parameters and
the result is a table (and, with --out, a JSON file) of the per-step records.

    python demo/pendulum_example.py                       # the reference's configuration
    python demo/pendulum_example.py --config my.yaml --steps 50 --cases DuSt-MPC DISCO
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributions as dist  # noqa: E402

from demo import configs  # noqa: E402

CASES = ("DuSt-MPC", "SVMPC", "MPPI Baseline", "DISCO")


def inst_cost(states, controls=None, n_pol=1, debug=None):
    theta, theta_d = states.chunk(2, dim=1)   # theta may range beyond 2 pi
    return 50.0 * (theta.cos() - 1) ** 2 + 1.0 * theta_d ** 2


def term_cost(states, n_pol=1, debug=None):
    return inst_cost(states).squeeze()


def run(cfg, cases=CASES, steps=None, episodes=None, seed=0, verbose=False):
    """-> pandas.DataFrame with the reference's columns plus `Case`."""
    import pandas as pd

    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import GaussianLikelihood
    from dust_b200.inference.mpf import MPF
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.kernels.base_kernels import RBF, RBFKernel
    from dust_b200.kernels.composite_kernels import iid_mp
    from dust_b200.models.pendulum import PendulumModel
    from dust_b200.utils.simulations import run_pendulum_simulation
    from dust_b200.utils.utf import MerweScaledUTF

    torch.manual_seed(seed)
    ep, sim = cfg["exp_params"], dict(cfg["sim_params"])
    sim["render"] = False
    if steps is not None:
        sim["steps"] = steps
    if episodes is not None:
        sim["episodes"] = episodes
    sim["verbose"] = verbose or sim.get("verbose", False)
    H, N, S, A = ep["horizon"], ep["n_particles"], ep["action_samples"], ep["ctrl_dim"]
    env_model = PendulumModel()
    init_state = torch.as_tensor(ep["init_state"]).clone()
    policies_prior = get_gmm(torch.randn(N, H, env_model.action_space.dim), torch.ones(N), ep["prior_sigma"] ** 2 * torch.eye(A))
    init_policies = policies_prior.sample([N])
    # the belief over (length, mass) the reference's script ends up using (its GMM over
    # params_prior_loc is overwritten by this box, pendulum_example.py:76-86)
    dynamics_prior = dist.Independent(dist.Uniform(torch.tensor([0.6, 0.6]), torch.tensor([1.3, 1.3])), 1)
    controller_kwargs = dict(observation_space=env_model.observation_space, action_space=env_model.action_space, hz_len=H,
                             action_samples=S, params_samples=ep["params_samples"], temperature=1 / ep["alpha"],
                             a_cov=ep["ctrl_sigma"] ** 2 * torch.eye(A), inst_cost_fn=inst_cost, term_cost_fn=term_cost)
    if ep["kernel"] == "message_passing":
        kernel = iid_mp(base_kernel=RBF(bandwidth=-1), ctrl_dim=1, indep_controls=True)
    elif ep["kernel"] == "rbf":
        kernel = RBFKernel()
    else:
        raise ValueError("Kernel type '{}' is not valid.".format(ep["kernel"]))
    lik_kwargs = {"alpha": ep["alpha"], "n_samples": S}
    svmpc_kwargs = dict(init_particles=init_policies, prior=policies_prior, kernel=kernel, n_particles=N,
                        bw_scale=ep["bandwidth_scaling"], n_steps=1, optimizer_class=torch.optim.SGD, lr=ep["learning_rate"])
    mpf_init = dynamics_prior.sample([ep["mpf_n_particles"]])
    if ep["mpf_log_space"]:
        mpf_init = mpf_init.clamp(min=1e-6).log()
    dynamics_lik = GaussianLikelihood(initial_obs=init_state, obs_std=ep["mpf_obs_std"],
                                      model=PendulumModel(uncertain_params=("length", "mass")), log_space=ep["mpf_log_space"])
    mpf = MPF(init_particles=mpf_init, likelihood=dynamics_lik, optimizer_class=torch.optim.SGD, lr=ep["mpf_learning_rate"],
              bw=ep["mpf_bandwidth"], bw_scale=ep["mpf_bandwidth_scaling"])
    tf = MerweScaledUTF(n=cfg["utf"]["n"], alpha=cfg["utf"]["alpha"])
    parameters_set = [{"length": float(item[0]), "mass": float(item[1])} for item in dynamics_prior.sample([sim["episodes"]])]
    common = dict(init_state=init_state, dyn_dist=dynamics_prior, experiment_params=parameters_set, **sim)
    frames = []
    for case in cases:
        print("\nRunning {} simulation:".format(case))
        if case == "DuSt-MPC":
            ctrl = MultiDISCO(params_sampling=True, n_policies=N, params_log_space=ep["mpf_log_space"], **controller_kwargs)
            df = run_pendulum_simulation(init_policies=init_policies, model_kwargs={"uncertain_params": ("length", "mass")},
                                         controller=ctrl, use_exact_model=False, use_svmpc=True, svmpc_kwargs=svmpc_kwargs,
                                         lik_kwargs=lik_kwargs, mpf=mpf, mpf_bw=ep["mpf_bandwidth"], **common)
        elif case == "SVMPC":
            ctrl = MultiDISCO(params_sampling=None, n_policies=N, **controller_kwargs)
            df = run_pendulum_simulation(init_policies=init_policies, model_kwargs={"uncertain_params": None}, controller=ctrl,
                                         use_exact_model=False, use_svmpc=True, svmpc_kwargs=svmpc_kwargs,
                                         lik_kwargs=lik_kwargs, mpf=None, **common)
        elif case == "MPPI Baseline":
            ctrl = MultiDISCO(params_sampling=None, n_policies=1, **controller_kwargs)
            df = run_pendulum_simulation(init_policies=init_policies[0].unsqueeze(0), model_kwargs={"uncertain_params": None},
                                         controller=ctrl, use_exact_model=True, use_svmpc=False, **common)
        elif case == "DISCO":
            ctrl = MultiDISCO(params_sampling=tf, n_policies=1, params_log_space=False, **controller_kwargs)
            df = run_pendulum_simulation(init_policies=init_policies[0].unsqueeze(0),
                                         model_kwargs={"uncertain_params": ("length", "mass")}, controller=ctrl,
                                         use_exact_model=False, use_svmpc=False, **common)
        else:
            raise ValueError("unknown case {!r}; choose from {}".format(case, CASES))
        df["Case"] = case
        frames.append(df)
    return pd.concat(frames, axis=0)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default=None, help="yaml file with the schema of the reference's pendulum_config.yaml")
    ap.add_argument("--cases", nargs="+", default=list(CASES))
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--episodes", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--out", default=None, help="write the per-step records as JSON")
    args = ap.parse_args()
    df = run(configs.load(args.config, configs.PENDULUM), args.cases, args.steps, args.episodes, args.seed, args.verbose)
    summary = df.groupby("Case")["Cost"].agg(["mean", "last"]).rename(columns={"mean": "mean step cost", "last": "final step cost"})
    print("\n" + summary.to_string())
    if args.out:
        with open(args.out, "w") as f:
            json.dump(df.reset_index().to_dict(), f)


if __name__ == "__main__":
    main()
