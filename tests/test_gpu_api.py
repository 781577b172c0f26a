"""The reference-shaped classes (MultiDISCO / SVMPC / MPF / likelihoods / models) driven the way
the reference's demos drive them, on the GPU, against the goldens recorded from the reference."""
import copy

import numpy as np
import pytest
import torch
import torch.distributions as dist

from oracle import dust_oracle as O
from tests.util import RTOL_COST, RTOL_PHI, assert_close_to_reference, golden_grid, load, rel_elem, rel_max

pytestmark = pytest.mark.gpu

ENV = dict(dt=0.015, control_type="acceleration", noise_std=[0.1, 0.1], init_state=[-9.0, -9.0, 0, 0],
           target_state=[9.0, 9.0, 0, 0], can_crash=True, with_obstacle=True, deterministic=True,
           cost_params=dict(w_qpos=0.5, w_qvel=0.25, w_ctrl=0.2, w_obs=1.0e6, w_qpos_T=1.0e3, w_qvel_T=0.1),
           obst_preset="grid_4x4", obst_width=2.1, max_speed=5, max_accel=10, map_cell_size=0.1, map_size=[22, 22],
           map_type="direct")


class FixedParams:
    """params_dist stand-in returning recorded samples (disco.py:168-174 contract)."""

    def __init__(self, samples, event_shape):
        self.samples, self.event_shape = samples, event_shape

    def sample(self, shape):
        return self.samples

    def log_prob(self, x):
        return torch.zeros(x.shape[0])


def demo_inst_cost(states, controls=None, n_pol=1, debug=None):
    # verbatim semantics of the demo's cost callable: the registry must recognise it by probing
    theta, theta_d = states.chunk(2, dim=1)
    return 50.0 * (theta.cos() - 1) ** 2 + 1.0 * theta_d ** 2


def demo_term_cost(states, n_pol=1, debug=None):
    return demo_inst_cost(states).squeeze()


def build_pendulum(d, kernel):
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import ExponentiatedUtility
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBF, RBFKernel
    from dust_b200.kernels.composite_kernels import iid_mp
    from dust_b200.models.pendulum import PendulumModel

    model = PendulumModel(uncertain_params=("length", "mass"))
    N, H, A = d["t0_in_theta0"].shape
    S = d["t0_in_eps"].shape[0]
    ctrl = MultiDISCO(observation_space=model.observation_space, action_space=model.action_space, hz_len=H,
                      n_policies=N, action_samples=S, params_samples=8, temperature=1.0, a_cov=4.0 * torch.eye(A),
                      inst_cost_fn=demo_inst_cost, term_cost_fn=demo_term_cost, params_sampling=True)
    prior = get_gmm(d["t0_in_mu0"], torch.ones(N), 4.0 * torch.eye(A))
    k = RBFKernel() if kernel == "rbf" else iid_mp(base_kernel=RBF(bandwidth=-1), ctrl_dim=A, indep_controls=True)
    lik = ExponentiatedUtility(1.0, n_samples=S, controller=ctrl, model=model)
    sv = SVMPC(init_particles=d["t0_in_theta0"].clone(), prior=prior, likelihood=lik, kernel=k, n_particles=N,
               bw_scale=1.0, n_steps=1, optimizer_class=torch.optim.SGD, lr=2.0, weighted_prior=False)
    return model, ctrl, sv


@pytest.mark.parametrize("name,kernel", [("svmpc_pendulum_rbf", "rbf"), ("svmpc_pendulum_mp", "mp")])
def test_svmpc_closed_loop_free_running(name, kernel):
    """optimize -> forward repeatedly, feeding only the recorded noise / parameter draws and plant
    states: particles, weights and the selected policy follow the reference step after step."""
    d = load(name)
    model, ctrl, sv = build_pendulum(d, kernel)
    for t in range(int(d["n_steps"])):
        gi, go = (lambda k: d[f"t{t}_in_{k}"]), (lambda k: d[f"t{t}_out_{k}"])
        pd = FixedParams(gi("params"), torch.Size([2]))
        sv.optimize(gi("state"), pd, eps=gi("eps"))
        assert rel_elem(sv.likelihood.last_costs.cpu(), go("costs")) <= 2e-4  # free running: inputs drift by rounding
        theta1 = sv.theta.cpu().clone()
        a_seq, pw = sv.forward(gi("state"), pd)
        assert int(sv.i_star) == int(go("i_star"))
        assert rel_max(theta1, go("theta1")) <= (5e-3 if kernel == "rbf" else 2e-4)
        assert rel_max(a_seq.cpu(), go("a_seq")) <= (5e-3 if kernel == "rbf" else 2e-4)
        assert float((pw.cpu() - go("p_weights")).abs().max()) <= 1e-3
        assert sv.likelihood.last_states.shape == (8, d["t0_in_eps"].shape[0], 3, 31, 2)
    # the prior exposed as a torch.distributions object is centred on the rolled particles
    assert torch.equal(sv.prior.component_distribution.base_dist.loc, sv.theta)


def test_multidisco_forward_and_step_api():
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.models.particle import Particle

    d = load("fwd_particle_s0")
    model = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
    S, N, H, A = d["actions"].shape
    ctrl = MultiDISCO(model.observation_space, model.action_space, H, N, S, temperature=float(d["temp"]),
                      a_cov=25.0 * torch.eye(A), params_sampling=True, params_samples=4, params_log_space=True,
                      inst_cost_fn=model.default_inst_cost, term_cost_fn=model.default_term_cost)
    ctrl.a_mat = d["a_mat0"].clone().cuda()
    pd = FixedParams(d["params"], torch.Size([1]))
    costs, states, actions, weights, plogp = ctrl.forward(d["state"], model, pd, d["actions"])
    assert costs.shape == (S, N) and states.shape == (4, S, N, H + 1, 4) and actions.shape == (4, S, N, H, A)
    assert rel_elem(costs.cpu(), d["costs"]) <= RTOL_COST
    assert torch.equal(states[:, :8].cpu(), d["states_sub"])
    assert rel_max(ctrl.a_mat.cpu(), d["a_mat1"]) <= 1e-4
    c2 = copy.deepcopy(ctrl)
    nxt = c2.step(strategy="argmax")
    assert rel_max(nxt.cpu(), d["step_argmax_action"]) <= 1e-4
    assert rel_max(c2.a_mat.cpu(), d["step_argmax_a_mat"]) <= 1e-4
    with pytest.raises(ValueError):
        ctrl.step(strategy="bogus")
    # internal sampling path runs and updates the plan
    ctrl.forward(d["state"], model, pd)
    assert torch.isfinite(ctrl.a_mat).all()


@pytest.mark.parametrize("kind", ["pendulum", "particle"])
def test_multidisco_control_regulariser(kind):
    """ctrl_penalty != 1 (disco.py:90, 334-344) in the stand-alone controller, two forward + step rounds
    against the reference's recording: costs, weights, plan update, mixture and the applied action."""
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.models.particle import Particle
    from dust_b200.models.pendulum import PendulumModel

    d = load(f"ctrlpen_{kind}")
    S, N, H, A = d["actions0"].shape
    if kind == "pendulum":
        model = PendulumModel(uncertain_params=("length", "mass"))
        cost = dict(inst_cost_fn=demo_inst_cost, term_cost_fn=demo_term_cost)
        ev = torch.Size([2])
    else:
        model = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
        cost = dict(inst_cost_fn=model.default_inst_cost, term_cost_fn=model.default_term_cost)
        ev = torch.Size([1])
    ctrl = MultiDISCO(model.observation_space, model.action_space, H, N, S, temperature=float(d["temp"]),
                      ctrl_penalty=float(d["ctrl_penalty"]), a_cov=torch.diag(d["sigma"] ** 2), params_sampling=True,
                      params_samples=d["params0"].shape[0], params_log_space=bool(d["log_space"]), **cost)
    assert abs(ctrl.a_reg - float(d["temp"]) * (1 - float(d["ctrl_penalty"]))) < 1e-7
    ctrl.a_mat = d["a_mat0"].clone().cuda()
    for it in range(2):
        assert rel_max(ctrl.a_seq.cpu(), d[f"a_seq_in{it}"]) <= 1e-4 or float(d[f"a_seq_in{it}"].abs().max()) == 0.0
        pd = FixedParams(d[f"params{it}"], ev)
        costs, states, actions, weights, _ = ctrl.forward(d["state"], model, pd, d[f"actions{it}"])
        assert rel_elem(costs.cpu(), d[f"costs{it}"]) <= (RTOL_COST if it == 0 else 1e-4)   # second round: inputs drift by rounding
        # soft-min weights against the oracle's reduction of the DEVICE costs (exact arithmetic check) ...
        w_o, delta_o, mix_o = O.softmin_update(costs.cpu().double(), (d[f"actions{it}"] - d[f"a_seq_in{it}"]).double(),
                                               float(d["temp"]))
        assert float((weights.cpu() - w_o).abs().max()) <= 1e-5
        assert rel_max(ctrl.a_mat.cpu(), d[f"a_mat_in{it}"].double() + delta_o) <= 1e-5
        assert float((ctrl.a_mix.cpu() - mix_o).abs().max()) <= 1e-5
        # ... and against the reference's own numbers where those are not rounding noise: with the particle
        # costs (4e7 inside an obstacle, float32 ulp 4) a cost ulp moves exp(-cost) by e^4, in the reference too
        if kind == "pendulum":
            assert float((weights.cpu() - d[f"weights{it}"]).abs().max()) <= 2e-3
            assert rel_max(ctrl.a_mat.cpu(), d[f"a_mat_fwd{it}"]) <= 2e-3
            assert rel_max(ctrl.a_mix.cpu(), d[f"a_mix{it}"]) <= 2e-3
        nxt = ctrl.step(strategy="average")
        if kind == "pendulum":
            assert rel_max(nxt.cpu(), d[f"action{it}"]) <= 2e-3
        # keep the second round on the reference's inputs (teacher forcing), so it checks the round itself
        if it == 0:
            ctrl.a_mat = d["a_mat_in1"].clone().cuda()
            ctrl.a_seq = d["a_seq_in1"].clone().cuda()


@pytest.mark.parametrize("name", ["utf_pendulum_n1_gmm", "utf_pendulum_n3_mvn"])
def test_multidisco_sigma_point_rollouts(name):
    """params_sampling = MerweScaledUTF (the demo's "DISCO" case, pendulum_example.py:143-147, 240-258):
    sigma points of the parameter belief rolled out on the device, the reference's weighted costs
    (disco.py:312-323, grouping quirk included), its state row order and its return shapes."""
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.models.pendulum import PendulumModel
    from dust_b200.utils.utf import MerweScaledUTF

    d = load(name)
    S, N, H, A = d["actions"].shape
    model = PendulumModel(uncertain_params=("length", "mass"))
    tf = MerweScaledUTF(n=2, alpha=0.5)
    ctrl = MultiDISCO(model.observation_space, model.action_space, H, N, S, temperature=float(d["temp"]),
                      a_cov=torch.diag(d["sigma"] ** 2), inst_cost_fn=demo_inst_cost, term_cost_fn=demo_term_cost,
                      params_sampling=tf, params_log_space=False)
    assert ctrl.n_params == 1 and ctrl.n_rollouts == S * N
    ctrl.a_mat = d["a_mat0"].clone().cuda()
    if bool(d["belief_is_mvn"]):
        pd = dist.MultivariateNormal(d["mean"], d["cov"])
    else:
        comp = dist.Independent(dist.Normal(d["locs"], float(d["comp_sigma"])), 1)
        pd = dist.MixtureSameFamily(dist.Categorical(torch.ones(d["locs"].shape[0])), comp)
    costs, states, actions, weights, plogp = ctrl.forward(d["state"], model, pd, d["actions"])
    assert tuple(actions.shape) == tuple(int(v) for v in d["acts_shape"])
    assert states.shape == d["states"].shape
    # trajectories: the device trig differs from libm by ~1e-7, states are compared to that
    assert float((states.cpu() - d["states"]).abs().max()) <= 2e-4
    assert rel_elem(costs.cpu(), d["costs"]) <= RTOL_COST
    assert float((weights.cpu() - d["weights"]).abs().max()) <= 5e-3
    assert rel_max(plogp.cpu(), d["params_log_p"]) <= 1e-5
    assert rel_max(ctrl.a_mat.cpu(), d["a_mat1"]) <= 5e-3
    nxt = copy.deepcopy(ctrl).step(strategy="average")
    assert rel_max(nxt.cpu(), d["action_avg"]) <= 5e-3
    with pytest.raises(NotImplementedError):
        MultiDISCO(model.observation_space, model.action_space, H, N, S, ctrl_penalty=0.5, params_sampling=tf,
                   inst_cost_fn=demo_inst_cost, term_cost_fn=demo_term_cost)


def test_model_step_and_costs_api():
    from dust_b200.models.particle import Particle
    from dust_b200.models.pendulum import PendulumModel

    cfg = O.ParticleCfg(golden_grid())
    torch.manual_seed(0)
    model = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
    x = torch.randn(500, 4) * torch.tensor([7.0, 7.0, 2.0, 2.0])
    a = torch.randn(500, 2) * 8
    m = torch.rand(500, 1) + 1.0
    assert torch.equal(model.step(x, a, {"mass": m}).cpu(), O.particle_step(cfg, x, a, m))
    assert torch.equal(model.step(x, a).cpu(), O.particle_step(cfg, x, a, 2.0))
    assert rel_max(model.default_inst_cost(x, a).cpu(), O.particle_inst_cost(cfg, x, a)) <= 1e-6
    assert rel_max(model.default_term_cost(x).cpu(), O.particle_term_cost(cfg, x)) <= 1e-6
    pend = PendulumModel(uncertain_params=("length", "mass"))
    xs, us = torch.randn(300, 2) * 3, torch.randn(300, 1) * 3
    lm = torch.rand(300, 2) * 0.7 + 0.6
    ref = O.pendulum_step(xs, us, length=lm[:, :1], mass=lm[:, 1:])
    assert rel_max(pend.step(xs, us, {"length": lm[:, :1], "mass": lm[:, 1:]}).cpu(), ref) <= 1e-6
    assert rel_max(pend.step(xs, us).cpu(), O.pendulum_step(xs, us)) <= 1e-6
    # very large angles leave the fast trig range and must still be right
    big = torch.tensor([[1.0e4, 0.5], [-3.0e6, -1.0], [70.0, 0.0]])
    assert rel_max(pend.step(big, torch.zeros(3, 1)).cpu(), O.pendulum_step(big, torch.zeros(3, 1))) <= 1e-6


def test_mpf_class_matches_reference():
    from dust_b200.inference.likelihoods import GaussianLikelihood
    from dust_b200.inference.mpf import MPF
    from dust_b200.models.particle import Particle

    d = load("dual_particle")
    model = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
    lik = GaussianLikelihood(initial_obs=d["t0_in_state"], obs_std=float(d["obs_std"]), model=model, log_space=True)
    mpf = MPF(init_particles=d["t0_in_mpf_x0"].clone(), likelihood=lik, optimizer_class=torch.optim.SGD,
              lr=float(d["mpf_lr"]), bw=float(d["mpf_prior_bw0"]), bw_scale=1.0)
    prior0 = mpf.prior  # captured once, as the demos do (particle_example.py:171)
    for t in range(int(d["n_steps"])):
        gn, bw = mpf.optimize(d[f"t{t}_out_a_seq"][0], d[f"t{t}_out_next_state"], bw=0.5, n_steps=20)
        assert rel_max(mpf.x.cpu(), d[f"t{t}_out_mpf_x1"]) <= 5e-4  # free running over 4 x 20 steps
        assert gn.shape == (20,)
    # the stale prior object tracks the particles through its aliased centres
    assert torch.equal(prior0.component_distribution.base_dist.loc, mpf.x)
    s = prior0.sample([4])
    assert s.shape == (4, 1) and s.is_cuda


def test_unsupported_paths_fail_loudly():
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.models.pendulum import PendulumModel

    model = PendulumModel()
    ctrl = MultiDISCO(model.observation_space, model.action_space, 5, 2, 4, inst_cost_fn=lambda s, *a, **k: s.sum(-1),
                      params_sampling=None)
    with pytest.raises(NotImplementedError):
        ctrl.forward(torch.zeros(2), model, None, torch.zeros(4, 2, 5, 1))
    # the control regulariser is supported stand-alone (test_multidisco_control_regulariser), not under SVMPC
    reg = MultiDISCO(model.observation_space, model.action_space, 5, 2, 4, inst_cost_fn=demo_inst_cost, ctrl_penalty=0.5)
    assert reg.a_reg == 0.5
    from dust_b200.inference.likelihoods import ExponentiatedUtility
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBFKernel
    with pytest.raises(NotImplementedError):
        SVMPC(init_particles=torch.zeros(2, 5, 1), prior=get_gmm(torch.zeros(2, 5, 1), torch.ones(2), torch.eye(1)),
              likelihood=ExponentiatedUtility(1.0, n_samples=4, controller=reg, model=model), kernel=RBFKernel(), n_particles=2,
              n_steps=1, optimizer_class=torch.optim.SGD, lr=1.0)
    with pytest.raises(ValueError):
        MultiDISCO(model.observation_space, model.action_space, 5, 2, 4, inst_cost_fn=demo_inst_cost,
                   params_sampling="sometimes")
    with pytest.raises(ValueError):
        MultiDISCO(model.observation_space, model.action_space, 5, 2, 4)


def test_batched_svmpc_equals_per_instance_runs():
    """BatchedSVMPC (B instances, one launch per stage) == B independent single-instance cores."""
    from dust_b200.batched import BatchedSVMPC
    from dust_b200.inference.core import SvmpcCore
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    B = 80
    ctl = BatchedSVMPC(PendulumModel(), B, 8, 64, 20, 2.0, 2.0, alpha=1.0, learning_rate=2.0, inst_cost_fn=inst_cost,
                       term_cost_fn=term_cost, seed=3)
    state = torch.randn(B, 2, device="cuda")
    eps = ctl.draw_noise().clone()
    theta0, mu0 = ctl.core.theta.clone(), ctl.core.mu.clone()
    ctl.optimize(state, eps)
    theta1 = ctl.core.theta.clone()
    a_seq, pw, i_star = ctl.forward()
    for b in (0, 41, B - 1):
        one = SvmpcCore(ctl.spec, theta0[b:b + 1].contiguous(), mu0[b:b + 1].contiguous(), torch.ones(1, 8, device="cuda"),
                        torch.tensor([4.0]), torch.tensor([2.0]), alpha=1.0, lr=2.0, kernel="gpytorch")
        one.optimize_step(state[b:b + 1].contiguous(), eps[b:b + 1].contiguous())
        assert rel_max(theta1[b].cpu(), one.theta[0].cpu()) <= 1e-5
        a1, p1, i1 = one.forward_step()
        assert int(i1[0]) == int(i_star[b])
        assert rel_max(a_seq[b].cpu(), a1[0].cpu()) <= 1e-5
        # and against the oracle
        st = O.SvmpcState(theta0[b].cpu().double(), mu0[b].cpu().double(), torch.ones(8).double(), 4.0)
        ref = O.svmpc_optimize(O.Model("pendulum"), st, state[b].cpu().double(), eps[b].cpu().double(),
                               torch.tensor([2.0]).double(), None, False, 1.0, 2.0, kernel="rbf")
        assert rel_max(theta1[b].cpu(), ref["theta1"]) <= RTOL_PHI


@pytest.mark.parametrize("kind,N,S,H,wp", [("pendulum", 8, 64, 20, False), ("pendulum", 3, 40, 30, True), ("particle", 5, 33, 12, True)])
def test_one_launch_control_step_matches_staged_path(kind, N, S, H, wp):
    """dust_svmpc_step (everything in the fused kernel's tail) == optimize_step + forward_step,
    first with a free-standing prior, then with the prior aliasing the particles."""
    from dust_b200 import _lib as L
    from dust_b200.inference.core import SvmpcCore
    from dust_b200.models.particle import Particle
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    torch.manual_seed(N * 7 + S)
    B = 90
    if kind == "pendulum":
        spec = PendulumModel(uncertain_params=("length", "mass")).device_spec(inst_cost, term_cost, "cuda")
        ds, A, dp = 2, 1, 2
    else:
        m = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
        spec = m.device_spec(m.default_inst_cost, m.default_term_cost, "cuda")
        ds, A, dp = 4, 2, 1
    dev = "cuda"
    state = torch.randn(B, ds, device=dev) * (torch.tensor([6.0, 6.0, 1.0, 1.0], device=dev) if kind == "particle" else 1.5)
    theta, mu = torch.randn(B, N, H, A, device=dev) * 2, torch.randn(B, N, H, A, device=dev)
    mix = torch.rand(B, N, device=dev) + 0.1
    params = torch.rand(B, 2, dp, device=dev) + 0.8
    mk = lambda: SvmpcCore(spec, theta.clone(), mu.clone(), mix.clone(), torch.full((A,), 4.0), torch.full((A,), 2.0),  # noqa: E731
                           alpha=1.0, lr=0.5, kernel="gpytorch", weighted_prior=wp)
    fused, staged = mk(), mk()
    staged._fused_ok = False          # keep this one on the staged kernels (gmm -> rollout/cost -> phi -> forward)
    lib = L.load()
    for step in range(2):  # step 0: free-standing prior; step 1: aliased prior
        eps = torch.randn(B, S, N, H, A, device=dev)
        lib.dust_profiler_reset(); lib.dust_profiler_enable(1)
        a1, p1, i1 = fused.control_step(state, eps, params, want_phi=True)
        prof = L.profiler_report(); lib.dust_profiler_enable(0)
        assert list(prof) == ["svmpc_instance_kernel"] and prof["svmpc_instance_kernel"][0] == 1, prof
        staged.optimize_step(state, eps, params)
        phi2 = staged.last["phi"]
        a2, p2, i2 = staged.forward_step()
        assert torch.equal(fused.last["costs"], staged.last["costs"])
        assert rel_max(fused.last["phi"].cpu(), phi2.cpu()) <= 1e-4
        assert rel_max(fused.theta.cpu(), staged.theta.cpu()) <= 1e-5
        assert torch.equal(i1, i2)
        assert float((p1 - p2).abs().max()) <= 1e-4
        assert rel_max(a1.cpu(), a2.cpu()) <= 1e-5
        assert rel_max(fused.mix.cpu(), staged.mix.cpu()) <= 1e-4
        # identical inputs for the next step (the two paths round differently at the 1e-7 level)
        staged.theta = fused.theta.clone()
        staged.mu, staged.mix = staged.theta, fused.mix.clone()


class SeqParams:
    """params_dist stand-in replaying a recorded sequence of draws, one per `sample` call."""

    def __init__(self, draws, event_shape=torch.Size([])):
        self.draws, self.i, self.event_shape = draws, 0, event_shape

    def sample(self, shape):
        x = self.draws[self.i]
        self.i += 1
        return x

    def log_prob(self, x):
        return torch.zeros(x.shape[0])


def test_particle_episode_driver_matches_reference_driver():
    """dust_b200.utils.simulations.run_particle_episode against an episode run by the reference's OWN
    driver (dust/utils/simulations.py:197-260; fixture made by tests/golden/make_golden.py episode):
    8 closed-loop steps, warm-up of 2, plant mass +1 at step 2, message-passing kernel, log-space
    parameter draws.  Only the recorded noise and parameter draws are fed in."""
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import ExponentiatedUtility
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBF
    from dust_b200.kernels.composite_kernels import iid_mp
    from dust_b200.models.particle import Particle
    from dust_b200.utils.simulations import run_particle_episode

    d = load("episode_particle_mp")
    N, H, A = d["theta0"].shape
    S = d["eps"].shape[1]
    model = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
    ctrl = MultiDISCO(model.observation_space, model.action_space, H, N, S, temperature=1.0, a_cov=25.0 * torch.eye(A),
                      params_sampling=True, params_samples=2, params_log_space=True,
                      inst_cost_fn=model.default_inst_cost, term_cost_fn=model.default_term_cost)
    prior = get_gmm(d["mu0"], torch.ones(N), 25.0 * torch.eye(A))
    lik = ExponentiatedUtility(1.0, n_samples=S, controller=ctrl, model=model)
    sv = SVMPC(init_particles=d["theta0"].clone(), prior=prior, likelihood=lik,
               kernel=iid_mp(base_kernel=RBF(bandwidth=-1), ctrl_dim=A, indep_controls=True), n_particles=N, bw_scale=1.0,
               n_steps=1, optimizer_class=torch.optim.SGD, lr=100.0, weighted_prior=True)
    draws = iter(d["eps"])
    lik.noise_fn = lambda shape: next(draws)
    hist = {}
    cum = run_particle_episode(d["init_state"], model, SeqParams(list(d["params"])), ctrl, use_svmpc=True,
                               warm_up=int(d["warm_up"]), svmpc=sv, load=float(d["load"]), steps=int(d["steps"]), history=hist)
    assert hist["states"].shape == d["plant_states"].shape
    # The loop is free running with lr = 100 and sharply peaked soft-min weights: a 1e-4 relative
    # difference in phi (the parity tolerance) moves the first controlled state by ~2e-6 and then
    # roughly doubles every step -- for the reference as much as for this build.  Tight where the
    # comparison is meaningful (warm-up and the first controlled steps), loose on the tail.
    per_step = (hist["states"].cpu() - d["plant_states"]).abs().max(1).values
    scale = float(d["plant_states"].abs().max())
    assert float(per_step[:2].max()) == 0.0, per_step.tolist()
    assert float(per_step[:5].max()) <= 5e-6 * scale, per_step.tolist()
    assert float(per_step.max()) <= 2e-3 * scale, per_step.tolist()
    assert abs(float(cum) - float(d["cum_cost"])) <= 1e-3 * float(d["cum_cost"]), (float(cum), float(d["cum_cost"]))
    # the policy particles themselves are chaotic over 8 free-running steps at lr = 100 (0.2 relative
    # apart by the end, in either implementation): only their shape and finiteness are asserted
    assert sv.theta.shape == d["theta_end"].shape and bool(torch.isfinite(sv.theta).all())
    assert torch.equal(hist["actions"][:2].cpu(), torch.zeros(2, A))


def test_pendulum_simulation_driver_runs_the_dual_loop():
    """run_pendulum_simulation (simulations.py:13-195) with SVMPC + MPF on the device: two short
    episodes; the frame has the reference's columns, finite costs, and the dynamics particles move."""
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import GaussianLikelihood
    from dust_b200.inference.mpf import MPF
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.kernels.base_kernels import RBFKernel
    from dust_b200.models.pendulum import PendulumModel
    from dust_b200.utils.simulations import run_pendulum_simulation

    torch.manual_seed(0)
    N, H, S, A = 3, 15, 32, 1
    model_kwargs = dict(uncertain_params=("length", "mass"))
    base_model = PendulumModel(**model_kwargs)
    ctrl = MultiDISCO(base_model.observation_space, base_model.action_space, H, N, S, temperature=1.0, a_cov=4.0 * torch.eye(A),
                      params_sampling=True, params_samples=4, inst_cost_fn=demo_inst_cost, term_cost_fn=demo_term_cost)
    mu0 = torch.randn(N, H, A)
    prior = get_gmm(mu0, torch.ones(N), 4.0 * torch.eye(A))
    init_policies = prior.sample([N])
    dyn_prior = dist.Uniform(torch.tensor([0.6, 0.6]), torch.tensor([1.3, 1.3]))
    init_state = torch.tensor([3.0, 0.0])
    mpf_lik = GaussianLikelihood(initial_obs=init_state, obs_std=0.1, model=base_model, log_space=False)
    mpf = MPF(init_particles=dyn_prior.sample([30]), likelihood=mpf_lik, optimizer_class=torch.optim.SGD, lr=1e-3, bw=0.05,
              bw_scale=1.0)
    x0 = mpf.x.clone()
    df = run_pendulum_simulation(init_state, init_policies, model_kwargs, mpf.prior,
                                 [dict(length=1.1, mass=0.8), dict(length=0.7, mass=1.2)], ctrl, use_svmpc=True,
                                 svmpc_kwargs=dict(init_particles=init_policies.clone(), prior=prior, kernel=RBFKernel(),
                                                   n_particles=N, bw_scale=1.0, n_steps=1, optimizer_class=torch.optim.SGD, lr=2.0),
                                 lik_kwargs=dict(alpha=1.0, n_samples=S), mpf=mpf, mpf_bw=0.05, mpf_steps=5, episodes=2, steps=6,
                                 warm_up=1)
    for col in ("Cost", "Position", "Speed", "Actions", "Timestep", "Iteration", "DynParticles", "DynBandwidths",
                "PolParticles", "Weights", "ExpParams", "AvgCumCost"):
        assert col in df.columns, col
    assert len(df) == 12 and set(df["Iteration"]) == {0, 1}
    assert np.isfinite(df["Cost"].to_numpy(dtype=float)).all()
    assert abs(float(df["Actions"].iloc[0])) == 0.0 and abs(float(df["Actions"].iloc[1])) > 0.0
    assert torch.equal(mpf.x, x0)          # the driver works on deep copies (simulations.py:62,78)
    moved = torch.tensor(df["DynParticles"].iloc[5])
    assert float((moved - x0.cpu()).abs().max()) > 0


def test_demo_runners_short_episodes():
    """demo/pendulum_example.py (all four cases, incl. the sigma-point DISCO baseline) and
    demo/particle_example.py (SVMPC + parameter filter) for a few control steps on their default
    configurations: finite records with the reference's columns; the swing-up cost does not blow up."""
    import math

    from demo import configs, particle_example, pendulum_example

    df = pendulum_example.run(configs.load(None, configs.PENDULUM), steps=12, episodes=1, seed=3)
    assert set(df["Case"]) == set(pendulum_example.CASES) and len(df) == 4 * 12
    for col in ("Cost", "Position", "Speed", "Actions", "AvgCumCost"):
        assert all(math.isfinite(v) for v in df[col]), col
    # MultiDISCO.step clips the plan to the action space (disco.py:408-410); SVMPC's best particle is logged as is
    assert df[df["Case"].isin(["MPPI Baseline", "DISCO"])]["Actions"].abs().max() <= 2.0 + 1e-6
    dual = df[df["Case"] == "DuSt-MPC"]
    assert all(len(p) == 50 for p in dual["DynParticles"])          # the filter's particles are logged each step
    cfg = configs.load(None, configs.PARTICLE)
    res = particle_example.run(cfg, steps=9, episodes=1, seed=1)
    r = res[0]
    assert r["steps"] == 9 and math.isfinite(r["cum_cost"])
    assert 1.0 < r["mass_estimate"] < 4.0
    assert all(math.isfinite(a) for row in r["actions"] for a in row)   # SVMPC's best particle is applied as is; the plant clips
    # warm-up steps apply the zero action (particle_example.py:182-184)
    assert all(a == 0.0 for row in r["actions"][: cfg["sim_params"]["warm_up"]] for a in row)
