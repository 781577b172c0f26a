"""Episode drivers with the call order of the reference's `dust/utils/simulations.py` (the
contract SURVEY.md section 3 describes), driving the device-backed classes of this package.

`run_particle_episode` keeps the reference signature (simulations.py:197-260) and adds the optional
dual-estimation arguments of the demo's own loop (demo/particle_example.py:177-207: `mpf`, `mpf_bw`,
`mpf_steps`).  `run_pendulum_simulation` keeps the signature of simulations.py:13-195; the reference
steps a gym `Pendulum-v0` as the plant, which this image does not have, so the plant here is
`PendulumModel.step` with the episode's true parameters (the controller-side code path is the same).
Rendering is not provided (plotting is outside the hot path)."""
from copy import deepcopy

import torch

from ..inference.likelihoods import ExponentiatedUtility
from ..inference.svmpc import SVMPC
from ..models.pendulum import PendulumModel


def run_particle_episode(init_state, model, dyn_dist, controller, use_svmpc=True, warm_up=30, svmpc: SVMPC = None, load=0,
                         steps=400, render=False, save_path=None, mpf=None, mpf_bw=None, mpf_steps=20, history=None):
    """Uses a copy of the controller's model as the simulated system, altering its load at steps//4.
    Returns the cumulative cost (inf after a crash), as the reference does.  `history`, if a dict, receives
    the visited states, actions and costs."""
    if render:
        raise NotImplementedError("rendering (matplotlib) is not part of dust_b200")
    system = deepcopy(model)
    dev = controller.device
    state = torch.as_tensor(init_state, dtype=torch.float32).to(dev)
    cum_cost = 0
    states_log, actions_log, costs_log = [], [], []
    for step in range(steps):
        if step == steps // 4:  # changes the simulator mass
            system.params_dict["mass"] = system.params_dict["mass"] + load
        if use_svmpc is True:
            svmpc.optimize(state, dyn_dist)
            if step < warm_up:
                action = torch.zeros(controller.dim_a, device=dev)
            else:
                a_seq, _ = svmpc.forward(state, dyn_dist)
                action = a_seq[0]
        else:
            controller.forward(state, model, params_dist=dyn_dist)
            action = controller.step(strategy="argmax")
        state = system.step(state, action.squeeze())
        if mpf is not None and step >= warm_up:
            mpf.optimize(action.squeeze(), state, bw=mpf_bw, n_steps=mpf_steps)   # updates mpf.prior in place
        cost = controller.inst_cost_fn(state.view(1, -1))
        cum_cost = cum_cost + cost
        states_log.append(state.detach().clone()); actions_log.append(action.detach().clone().reshape(-1)); costs_log.append(cost.detach().clone())
        if system.with_obstacle and bool(system.obst_map.get_collisions(state[:2].cpu())):
            print("Crashed at step {}".format(step))
            cum_cost = float("inf")
            break
        if float((system.target.to(state.device) - state).norm()) <= 1.0:
            break
    if isinstance(history, dict):
        history.update(states=torch.stack(states_log), actions=torch.stack(actions_log), costs=torch.stack(costs_log).reshape(-1))
    return cum_cost


def run_pendulum_simulation(init_state, init_policies, model_kwargs, dyn_dist, experiment_params, controller,
                            use_exact_model=True, use_svmpc=True, svmpc_kwargs=None, lik_kwargs=None, mpf=None, mpf_bw=None,
                            mpf_steps=20, episodes=3, steps=200, render=False, warm_up=1, verbose=False,
                            steps_per_message=20):
    """One DataFrame row per time step and episode, with the reference's columns."""
    import pandas as pd

    if render:
        raise NotImplementedError("rendering (gym viewer) is not part of dust_b200")
    epoch_df = pd.DataFrame()
    dev = controller.device
    for i in range(episodes):
        if use_exact_model:
            model = PendulumModel(**experiment_params[i], **model_kwargs)
        else:
            model = PendulumModel(length=float(dyn_dist.mean[0]), mass=float(dyn_dist.mean[1]), **model_kwargs)
        plant = PendulumModel(**experiment_params[i], **{k: v for k, v in model_kwargs.items() if k != "uncertain_params"})
        state = torch.as_tensor(init_state, dtype=torch.float32).to(dev).reshape(1, -1)
        sim_ctrl = deepcopy(controller)
        sim_ctrl.a_mat = torch.as_tensor(init_policies).detach().clone().to(dev)
        sim_svmpc = None
        if use_svmpc:
            assert svmpc_kwargs is not None and lik_kwargs is not None, \
                "Need a Stein Optimizer and likelihood for dual svmpc simulation."
            likelihood = ExponentiatedUtility(**lik_kwargs, controller=sim_ctrl, model=model)
            sim_svmpc = SVMPC(likelihood=likelihood, **svmpc_kwargs)
        sim_mpf, dyn_particles, dyn_bws, ep_dyn = None, None, None, dyn_dist
        if mpf is not None:
            sim_mpf = deepcopy(mpf)
            ep_dyn = sim_mpf.prior
            dyn_particles = torch.full((steps, *sim_mpf.x.shape), float("nan"))
            dyn_bws = torch.zeros(steps)
        nan = float("nan")
        states = torch.full((steps, sim_ctrl.dim_s), nan)
        actions = torch.full((steps, sim_ctrl.dim_a), nan)
        costs = torch.full((steps, 1), nan)
        pol_particles = torch.full((steps, sim_ctrl.n_pol, sim_ctrl.hz_len, sim_ctrl.dim_a), nan)
        weights = torch.full((steps, sim_ctrl.n_pol), nan)
        for step in range(steps):
            if use_svmpc:
                sim_svmpc.optimize(state, ep_dyn)
                if step < warm_up:
                    action = torch.zeros(sim_ctrl.dim_a, device=dev)
                else:
                    a_seq, p_weights = sim_svmpc.forward(state, ep_dyn)
                    action = a_seq[0]
                    pol_particles[step] = sim_svmpc.theta.detach().cpu()
                    weights[step] = p_weights.detach().cpu()
            else:
                sim_ctrl.forward(state, model, ep_dyn)
                action = sim_ctrl.step(strategy="average").flatten()
            actions[step] = action.detach().cpu()
            state = plant.step(state, action.reshape(1, -1)).reshape(1, -1)
            if sim_mpf is not None:
                _, bw = sim_mpf.optimize(action.squeeze(), state, bw=mpf_bw, n_steps=mpf_steps)
                dyn_particles[step] = sim_mpf.x.detach().cpu()
                dyn_bws[step] = float(bw)
            cost = sim_ctrl.inst_cost_fn(state.view(1, -1))
            if verbose and not step % steps_per_message:
                print("Step {0}: action taken {1:.2f}, cost {2:.2f}".format(step, float(action), float(cost)))
            states[step] = state.detach().cpu()
            costs[step] = cost.detach().cpu().reshape(-1)
        episode_df = pd.DataFrame(index=list(range(steps)), data={
            "Cost": costs[:, 0].tolist(), "Position": states[:, 0].tolist(), "Speed": states[:, 1].tolist(),
            "Actions": actions[:, 0].tolist(), "Timestep": list(range(steps)), "Iteration": i,
            "DynParticles": dyn_particles.tolist() if dyn_particles is not None else None,
            "DynBandwidths": dyn_bws.tolist() if dyn_bws is not None else None,
            "PolParticles": pol_particles[..., 0, 0].tolist(), "Weights": weights.tolist(),
            "ExpParams": steps * [list(experiment_params[i].values())]})
        episode_df["AvgCumCost"] = (episode_df["Cost"].cumsum(0) / (episode_df["Timestep"] + 1)).round(2)
        epoch_df = pd.concat((epoch_df, episode_df), axis=0)
    return epoch_df
