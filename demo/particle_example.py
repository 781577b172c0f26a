#!/usr/bin/env python
"""2-D point-mass navigation among obstacles with multi-policy SVMPC and, optionally, the dynamics-parameter
filter estimating the mass while the simulated system gains an extra load after a quarter of the episode
(the loop of the reference's demo/particle_example.py:150-254 without its plotting).

    python demo/particle_example.py --steps 400
    python demo/particle_example.py --config my.yaml --out episode.json
"""
import argparse
import json
import os
import sys
from copy import deepcopy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributions as dist  # noqa: E402

from demo import configs  # noqa: E402


def build(cfg, seed=0):
    """The objects of one experiment: (controller, svmpc, mpf | None, model, dynamics_prior)."""
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference import likelihoods
    from dust_b200.inference.mpf import MPF
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBF, RBFKernel
    from dust_b200.kernels.composite_kernels import iid_mp
    from dust_b200.models.particle import Particle

    torch.manual_seed(seed)
    ep, env = cfg["exp_params"], cfg["env_params"]
    H, N, S, A = ep["horizon"], ep["n_particles"], ep["action_samples"], ep["ctrl_dim"]
    state = torch.as_tensor(env["init_state"], dtype=torch.float).clone()
    policies_prior = get_gmm(torch.randn(N, H, A), torch.ones(N), ep["prior_sigma"] ** 2 * torch.eye(A))
    init_policies = policies_prior.sample([N])
    dynamics_prior = getattr(dist, ep["dyn_prior"])(ep["dyn_prior_arg1"], ep["dyn_prior_arg2"])
    model = Particle(**env, uncertain_params=["mass"], mass=dynamics_prior.mean)
    controller = MultiDISCO(model.observation_space, model.action_space, H, N, S, temperature=1 / ep["alpha"],
                            a_cov=ep["ctrl_sigma"] ** 2 * torch.eye(A), params_sampling=ep["sampling"],
                            params_samples=ep["params_samples"], params_log_space=ep["mpf_log_space"],
                            inst_cost_fn=model.default_inst_cost, term_cost_fn=model.default_term_cost)
    if ep["kernel"] == "message_passing":
        kernel = iid_mp(base_kernel=RBF(bandwidth=-1), ctrl_dim=2, indep_controls=True)
    elif ep["kernel"] == "rbf":
        kernel = RBFKernel()
    else:
        raise ValueError("Kernel type '{}' is not valid.".format(ep["kernel"]))
    lik = getattr(likelihoods, ep["likelihood"])(ep["alpha"], controller=controller, model=model, n_samples=S)
    svmpc = SVMPC(init_particles=init_policies.detach().clone(), prior=policies_prior, likelihood=lik, kernel=kernel,
                  n_particles=N, bw_scale=ep["bandwidth_scaling"], n_steps=1, optimizer_class=torch.optim.SGD,
                  lr=ep["learning_rate"], weighted_prior=ep["weighted_prior"])
    mpf = None
    if ep["use_mpf"]:
        mpf_init = dynamics_prior.sample([ep["mpf_n_particles"], 1]).clamp(min=1e-6)
        mpf_init = mpf_init.log() if ep["mpf_log_space"] else mpf_init
        dyn_lik = likelihoods.GaussianLikelihood(initial_obs=state, obs_std=ep["mpf_obs_std"], model=model,
                                                 log_space=ep["mpf_log_space"])
        mpf = MPF(init_particles=mpf_init, likelihood=dyn_lik, optimizer_class=torch.optim.SGD, lr=ep["mpf_learning_rate"],
                  bw=(2 * ep["dyn_prior_arg2"]) ** 1 / 2, bw_scale=ep["mpf_bandwidth_scaling"])   # sic: (2 s)**1 / 2 = s
    return controller, svmpc, mpf, model, dynamics_prior


def run(cfg, steps=None, episodes=None, seed=0):
    """-> list of per-episode dicts: cumulative cost, visited states, actions, step costs, mass estimates."""
    from dust_b200.utils.simulations import run_particle_episode

    ep, sim = cfg["exp_params"], cfg["sim_params"]
    steps = sim["steps"] if steps is None else steps
    episodes = sim["episodes"] if episodes is None else episodes
    base = build(cfg, seed)
    results = []
    for e in range(episodes):
        controller, svmpc, mpf, model, dynamics_prior = deepcopy(base)
        dyn = mpf.prior if mpf is not None else dynamics_prior
        hist = {}
        cum = run_particle_episode(torch.as_tensor(cfg["env_params"]["init_state"], dtype=torch.float), model, dyn, controller,
                                   use_svmpc=ep["use_svmpc"], warm_up=sim["warm_up"], svmpc=svmpc, load=ep["extra_load"],
                                   steps=steps, mpf=mpf, mpf_bw=ep["mpf_bandwidth"], mpf_steps=ep["mpf_steps"], history=hist)
        mass = None
        if mpf is not None:
            x = mpf.x.detach().cpu()
            mass = float((x.exp() if ep["mpf_log_space"] else x).mean())
        results.append(dict(episode=e, cum_cost=float(cum), steps=int(hist["states"].shape[0]), mass_estimate=mass,
                            states=hist["states"].cpu().tolist(), actions=hist["actions"].cpu().tolist(),
                            costs=hist["costs"].cpu().tolist()))
        print("episode {}: {} steps, cumulative cost {:.1f}, mass estimate {}".format(
            e, results[-1]["steps"], results[-1]["cum_cost"], "-" if mass is None else "{:.3f}".format(mass)))
    return results


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default=None, help="yaml file with the schema of the reference's particle_config.yaml")
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--episodes", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None, help="write the episode records as JSON")
    args = ap.parse_args()
    res = run(configs.load(args.config, configs.PARTICLE), args.steps, args.episodes, args.seed)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f)


if __name__ == "__main__":
    main()
