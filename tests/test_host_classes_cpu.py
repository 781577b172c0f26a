"""Host logic of the reference-shaped classes on CPU, first of all the non-SGD optimiser path of SVMPC (svgd.py:115: Adam is the reference's default; plain SGD is
what the demos use and what the update kernels fuse).  `SvmpcCore` is run HERE ON CPU with oracle-backed stand-ins
for the library's ops, against recordings of the unmodified reference stepping Adam / momentum SGD: this pins the
optimiser plumbing -- `grad = -phi`, in-place steps, two SVGD steps per control step, and the reference's quirk
that the optimiser state is keyed by the particle tensor and therefore starts afresh after every roll
(svmpc.py:144,158).  The device kernels behind the real ops have their own parity tests (`-m gpu`)."""
import functools

import pytest
import torch

from oracle import dust_oracle as O
from tests.util import load, rel_max

torch.set_num_threads(1)
MODEL = O.Model("pendulum")


class OracleOps:
    """CPU stand-ins with the signatures `SvmpcCore` uses (B = 1), computed in float64 by the oracle."""

    @staticmethod
    def gmm_log_norm(var_full):
        from dust_b200 import ops

        return ops.gmm_log_norm(var_full)

    @staticmethod
    def gmm(x, mu, mix, inv_var, log_norm, want_log_prob=True, want_score=True):
        var = 1.0 / inv_var.double()
        return None, O.gmm_score(x[0].double(), mu[0].double(), mix[0].double(), var).float().unsqueeze(0)

    @staticmethod
    def rollout_cost(spec, state0, eps, theta=None, sigma=None, params=None, alpha=1.0, want=(), **kw):
        actions = theta[0].double() + sigma.double() * eps[0].double()
        out = O.disco_forward(MODEL, state0[0].double(), actions, None if params is None else params[0].double())
        costs = out["costs"]
        res = {"costs": costs.float().unsqueeze(0), "log_lik": O.exp_utility_log_prob(costs, alpha).float().unsqueeze(0)}
        res["grad_lik"] = O.analytic_lik_grad(costs, actions, theta[0].double(), sigma.double(), alpha).float().unsqueeze(0)
        return res

    @staticmethod
    def svgd_phi(x, score, gamma=0.0, c1=0.0, c2=0.0, lr=0.0, want_update=False, per_dim=False, **kw):
        if per_dim:      # the message-passing kernel (composite_kernels.py:33-64, svmpc.py:64-74)
            phi = O.phi_svmpc_iid_mp(x[0].double(), score[0].double()).float().unsqueeze(0)
        else:
            phi = O.phi_unified(x[0].double(), score[0].double(), gamma, c1, c2).float().unsqueeze(0)
        return {"phi": phi, "x_out": (x + lr * phi) if want_update else None}

    @staticmethod
    def svmpc_step(*a, **k):
        raise NotImplementedError("no one-launch step in the stand-in")

    @staticmethod
    def svmpc_forward(log_lik, theta, mu, mix, inv_var, log_norm, roll_strategy=0, weighted_prior=False, resample_noise=None):
        N = theta.shape[1]
        flat = lambda t: t[0].reshape(N, -1).double()  # noqa: E731
        log_w = log_lik[0].double() + O.gmm_log_prob(flat(theta), flat(mu), mix[0].double(), 1.0 / inv_var.double())
        p = (log_w - log_w.logsumexp(0)).exp()
        i_star = int(p.argmax())
        th = theta[0].roll(-1, dims=-2).clone()
        th[..., -1, :] = th[..., -2, :]
        mix_next = p.float() if weighted_prior else torch.ones(N)
        return dict(a_seq=theta[0, i_star].clone().unsqueeze(0), p_weights=p.float().unsqueeze(0),
                    i_star=torch.tensor([i_star]), theta_next=th.unsqueeze(0), mix_next=mix_next.unsqueeze(0))


@pytest.mark.parametrize("name,opt", [("svmpc_pendulum_adam", functools.partial(torch.optim.Adam)),
                                      ("svmpc_pendulum_momentum", functools.partial(torch.optim.SGD, momentum=0.9))])
def test_svmpc_core_with_torch_optimizers_follows_the_reference(name, opt):
    from dust_b200.inference.core import SvmpcCore

    d = load(name)
    N = d["theta_init"].shape[0]
    core = SvmpcCore(None, d["theta_init"].clone().unsqueeze(0), d["mu_init"].clone().unsqueeze(0), torch.ones(1, N),
                     torch.tensor([float(d["prior_var"])]), d["sigma"], alpha=1.0, lr=float(d["lr"]), kernel="gpytorch",
                     optimizer=functools.partial(opt, lr=float(d["lr"])), ops_module=OracleOps)
    for t in range(int(d["n_ctrl"])):
        state0 = d[f"t{t}_state"].reshape(1, -1)
        first_opt = None
        for k in range(2):                      # n_steps = 2: two SVGD steps per control step, own draws each
            core.optimize_step(state0, d[f"t{t}_eps"][k].unsqueeze(0), d[f"t{t}_params"][k].unsqueeze(0))
            first_opt = first_opt or core._opt
            assert core._opt is first_opt        # one optimiser (and one state) within a control step ...
        assert rel_max(core.last["costs"][0], d[f"t{t}_costs"]) <= 1e-5
        assert rel_max(core.theta[0], d[f"t{t}_theta1"]) <= 2e-4
        a_seq, p_w, i_star = core.forward_step()
        assert int(i_star[0]) == int(d[f"t{t}_i_star"])
        assert rel_max(a_seq[0], d[f"t{t}_a_seq"]) <= 2e-4
        assert float((p_w[0] - d[f"t{t}_p_weights"]).abs().max()) <= 1e-3
        assert rel_max(core.theta[0], d[f"t{t}_theta2"]) <= 2e-4
    # ... and a new one after the roll replaced the particle tensor
    core.optimize_step(state0, d["t0_eps"][0].unsqueeze(0), d["t0_params"][0].unsqueeze(0))
    assert core._opt is not first_opt


def test_plain_sgd_keeps_the_fused_update():
    """optimizer=None: the update comes out of the phi kernel (x_out); the staged fallback of control_step is
    taken when the one-launch step declines."""
    from dust_b200.inference.core import SvmpcCore

    d = load("svmpc_pendulum_adam")
    N = d["theta_init"].shape[0]
    core = SvmpcCore(None, d["theta_init"].clone().unsqueeze(0), d["mu_init"].clone().unsqueeze(0), torch.ones(1, N),
                     torch.tensor([float(d["prior_var"])]), d["sigma"], alpha=1.0, lr=0.5, kernel="gpytorch", ops_module=OracleOps)
    theta0 = core.theta.clone()
    a_seq, p_w, i_star = core.control_step(d["t0_state"].reshape(1, -1), d["t0_eps"][0].unsqueeze(0), d["t0_params"][0].unsqueeze(0))
    assert core._opt is None and core._fused_ok is False
    assert a_seq.shape == (1, theta0.shape[2], theta0.shape[3]) and not torch.equal(core.theta, theta0)
    import copy

    assert copy.deepcopy(core).theta.shape == core.theta.shape      # the ops stand-in / module never blocks deepcopy


def test_weights_from_a_fresh_likelihood_sample():
    """`get_weights(fast_pred=False)` / `forward(fast_pred=False)` (svmpc.py:128-140, 172-200): new noise and
    parameter draws, rollouts at the UPDATED particles, weights from those costs -- against the reference's run
    (optimise, peek at the weights, forward; three draws per control step)."""
    import copy

    from dust_b200.inference.core import SvmpcCore

    d = load("svmpc_pendulum_slow_pred")
    N = d["theta_init"].shape[0]
    core = SvmpcCore(None, d["theta_init"].clone().unsqueeze(0), d["mu_init"].clone().unsqueeze(0), torch.ones(1, N),
                     torch.tensor([float(d["prior_var"])]), d["sigma"], alpha=1.0, lr=float(d["lr"]), kernel="gpytorch",
                     ops_module=OracleOps)
    for t in range(int(d["n_ctrl"])):
        state0 = d[f"t{t}_state"].reshape(1, -1)
        eps, params = d[f"t{t}_eps"], d[f"t{t}_params"]
        core.optimize_step(state0, eps[0].unsqueeze(0), params[0].unsqueeze(0))
        peek = copy.copy(core)                                                   # SVMPC._peek_weights
        pw_peek = peek.forward_step(peek.likelihood_at_particles(state0, eps[1].unsqueeze(0), params[1].unsqueeze(0)))[1][0]
        assert float((pw_peek - d[f"t{t}_peek"]).abs().max()) <= 2e-3
        log_lik = core.likelihood_at_particles(state0, eps[2].unsqueeze(0), params[2].unsqueeze(0))
        assert rel_max(core.last["costs"][0], d[f"t{t}_costs_fwd"]) <= 1e-4
        a_seq, p_w, _ = core.forward_step(log_lik)
        assert float((p_w[0] - d[f"t{t}_p_weights"]).abs().max()) <= 2e-3
        assert rel_max(a_seq[0], d[f"t{t}_a_seq"]) <= 5e-4
        assert rel_max(core.theta[0], d[f"t{t}_theta2"]) <= 5e-4


# ---- the drop-in SVMPC class itself, on CPU: device="cpu" objects + the oracle-backed ops ---------------------
class _FixedParams:
    """params_dist stand-in returning a recorded draw (disco.py:168-174 contract)."""

    def __init__(self, samples):
        self.samples, self.event_shape = samples, torch.Size([samples.shape[-1]])

    def sample(self, shape):
        return self.samples

    def log_prob(self, x):
        return torch.zeros(x.shape[0])


def _swingup_cost(states, controls=None, n_pol=1, debug=None):
    theta, theta_d = states.chunk(2, dim=1)
    return 50.0 * (theta.cos() - 1) ** 2 + 1.0 * theta_d ** 2


def _build_svmpc(d, **svmpc_kwargs):
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import ExponentiatedUtility
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBFKernel
    from dust_b200.models.pendulum import PendulumModel

    N, H, A = d["theta_init"].shape
    S = d["t0_eps"].shape[1]
    model = PendulumModel(uncertain_params=("length", "mass"))
    ctrl = MultiDISCO(observation_space=model.observation_space, action_space=model.action_space, hz_len=H, n_policies=N,
                      action_samples=S, params_samples=3, temperature=1.0, a_cov=4.0 * torch.eye(A), inst_cost_fn=_swingup_cost,
                      term_cost_fn=lambda s, **k: _swingup_cost(s).squeeze(), params_sampling=True, device="cpu")
    lik = ExponentiatedUtility(1.0, n_samples=S, controller=ctrl, model=model)
    sv = SVMPC(init_particles=d["theta_init"].clone(), prior=get_gmm(d["mu_init"], torch.ones(N), 4.0 * torch.eye(A)),
               likelihood=lik, kernel=RBFKernel(), n_particles=N, bw_scale=1.0, lr=float(d["lr"]), **svmpc_kwargs)
    sv._core._ops_module = OracleOps
    return sv


@pytest.mark.parametrize("name,kw", [("svmpc_pendulum_adam", dict(optimizer_class=torch.optim.Adam)),
                                     ("svmpc_pendulum_momentum", dict(optimizer_class=torch.optim.SGD, momentum=0.9))])
def test_svmpc_class_with_torch_optimizers(name, kw):
    """`SVMPC(optimizer_class=...)` as a user of the reference writes it, stepped through `step` / `forward`."""
    d = load(name)
    sv = _build_svmpc(d, n_steps=2, **kw)
    assert sv._core._make_opt is not None
    for t in range(int(d["n_ctrl"])):
        st = d[f"t{t}_state"]
        for k in range(2):
            sv.step(st, _FixedParams(d[f"t{t}_params"][k]), eps=d[f"t{t}_eps"][k])
        assert rel_max(sv.theta, d[f"t{t}_theta1"]) <= 2e-4
        a_seq, pw = sv.forward(st, None)
        assert int(sv.i_star) == int(d[f"t{t}_i_star"])
        assert rel_max(a_seq, d[f"t{t}_a_seq"]) <= 2e-4 and rel_max(sv.theta, d[f"t{t}_theta2"]) <= 2e-4
        assert torch.equal(sv.prior.component_distribution.base_dist.loc, sv.theta)      # prior refreshed on the rolled particles


def test_svmpc_class_slow_prediction_and_plain_sgd():
    import copy

    d = load("svmpc_pendulum_slow_pred")
    sv = _build_svmpc(d, n_steps=1, optimizer_class=torch.optim.SGD)
    assert sv._core._make_opt is None                                   # plain SGD: the fused update
    for t in range(int(d["n_ctrl"])):
        st, eps, prm = d[f"t{t}_state"], d[f"t{t}_eps"], d[f"t{t}_params"]
        sv.optimize(st, _FixedParams(prm[0]), eps=eps[0])
        theta1 = sv.theta.clone()
        pk = sv.get_weights(st, _FixedParams(prm[1]), fast_pred=False, eps=eps[1])
        assert torch.equal(sv.theta, theta1), "peeking at the weights must not roll the particles"
        assert float((pk - d[f"t{t}_peek"]).abs().max()) <= 2e-3
        a_seq, pw = sv.forward(st, _FixedParams(prm[2]), fast_pred=False, eps=eps[2])
        assert rel_max(sv.likelihood.last_costs, d[f"t{t}_costs_fwd"]) <= 1e-4
        assert float((pw - d[f"t{t}_p_weights"]).abs().max()) <= 2e-3
        assert rel_max(a_seq, d[f"t{t}_a_seq"]) <= 5e-4 and rel_max(sv.theta, d[f"t{t}_theta2"]) <= 5e-4
    assert copy.deepcopy(sv).theta.shape == sv.theta.shape             # the demos deep-copy the optimiser per episode
    with pytest.raises(NotImplementedError):
        sv.forward(d["t0_state"], None, steps=-2)


@pytest.mark.parametrize("name,kernel", [("svmpc_pendulum_rbf", "rbf"), ("svmpc_pendulum_mp", "mp")])
def test_svmpc_class_closed_loop_on_cpu(name, kernel):
    """The CPU twin of the GPU closed-loop test: the reference-shaped classes (MultiDISCO, ExponentiatedUtility,
    SVMPC) driven as the demos drive them, with the oracle standing in for the device ops -- so a break in the
    host layer shows up in the CPU suite, not only on a B200."""
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import ExponentiatedUtility
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBF, RBFKernel
    from dust_b200.kernels.composite_kernels import iid_mp
    from dust_b200.models.pendulum import PendulumModel

    d = load(name)
    N, H, A = d["t0_in_theta0"].shape
    S = d["t0_in_eps"].shape[0]
    model = PendulumModel(uncertain_params=("length", "mass"))
    ctrl = MultiDISCO(observation_space=model.observation_space, action_space=model.action_space, hz_len=H, n_policies=N,
                      action_samples=S, params_samples=8, temperature=1.0, a_cov=4.0 * torch.eye(A), inst_cost_fn=_swingup_cost,
                      term_cost_fn=lambda s, **k: _swingup_cost(s).squeeze(), params_sampling=True, device="cpu")
    k = RBFKernel() if kernel == "rbf" else iid_mp(base_kernel=RBF(bandwidth=-1), ctrl_dim=A, indep_controls=True)
    sv = SVMPC(init_particles=d["t0_in_theta0"].clone(), prior=get_gmm(d["t0_in_mu0"], torch.ones(N), 4.0 * torch.eye(A)),
               likelihood=ExponentiatedUtility(1.0, n_samples=S, controller=ctrl, model=model), kernel=k, n_particles=N,
               bw_scale=1.0, n_steps=1, optimizer_class=torch.optim.SGD, lr=2.0, weighted_prior=False)
    sv._core._ops_module = OracleOps
    for t in range(int(d["n_steps"])):
        gi, go = (lambda key: d[f"t{t}_in_{key}"]), (lambda key: d[f"t{t}_out_{key}"])
        sv.optimize(gi("state"), _FixedParams(gi("params")), eps=gi("eps"))
        assert rel_max(sv.likelihood.last_costs, go("costs")) <= 2e-4
        theta1 = sv.theta.clone()
        a_seq, pw = sv.forward(gi("state"), None)
        assert int(sv.i_star) == int(go("i_star"))
        assert rel_max(theta1, go("theta1")) <= (5e-3 if kernel == "rbf" else 2e-4)
        assert rel_max(a_seq, go("a_seq")) <= (5e-3 if kernel == "rbf" else 2e-4)
        assert float((pw - go("p_weights")).abs().max()) <= 1e-3


# ---- the stand-alone controller on CPU: dust_b200.controllers.disco.ops replaced by oracle-backed stand-ins ----
class OracleControllerOps:
    @staticmethod
    def rollout_cost(spec, state0, noise, theta=None, sigma=None, params=None, param_tiling=0, a_seq=None, pert=None,
                     temperature=1.0, want=(), sigma_weights=None, ctrl_mat=None, ctrl_reg=0.0, **kw):
        actions = (noise[0] if theta is None else theta[0] + sigma * noise[0]).double()
        aseq = None if a_seq is None else a_seq[0].double()
        if sigma_weights is not None:
            out = O.disco_forward_sigma(MODEL, state0[0].double(), actions, params[0].T.double(), sigma_weights.double(),
                                        temperature, aseq)
            S, N, H, A = actions.shape
            pts = sigma_weights.numel()
            states = out["states"].reshape(S, N, pts, H + 1, MODEL.ds).permute(2, 0, 1, 3, 4)     # the device layout [P,S,N,..]
        else:
            prm = None if params is None else (params[0].reshape(-1) if param_tiling == 1 else params[0]).double()
            out = O.disco_forward(MODEL, state0[0].double(), actions, prm, False, temperature, a_seq=aseq, a_reg=ctrl_reg,
                                  a_mat=None if ctrl_mat is None else ctrl_mat[0].double(), a_pre=torch.eye(actions.shape[-1]).double(),
                                  eps=None if pert is None else pert[0].double())
            states = out["states"]
        res = dict(costs=out["costs"], mppi_weights=out["weights"], mppi_delta=out["delta"], mix=out["a_mix"], states=states)
        return {k: v.float().unsqueeze(0) for k, v in res.items() if k in want}

    @staticmethod
    def disco_step(a_mat, a_mix, low, high, strategy=0, steps=1):
        nxt, a_seq, new_mat = O.disco_step(a_mat[0], a_mix[0], low, high, "argmax" if strategy == 0 else "average", steps)
        a_mat[0] = new_mat
        return nxt.unsqueeze(0), a_seq.unsqueeze(0)


def _controller(monkeypatch, d, S, N, H, A, **kw):
    from dust_b200.controllers import disco as disco_module
    from dust_b200.models.pendulum import PendulumModel

    monkeypatch.setattr(disco_module, "ops", OracleControllerOps)
    model = PendulumModel(uncertain_params=("length", "mass"))
    ctrl = disco_module.MultiDISCO(model.observation_space, model.action_space, H, N, S, temperature=float(d["temp"]),
                                   a_cov=torch.diag(d["sigma"] ** 2), inst_cost_fn=_swingup_cost,
                                   term_cost_fn=lambda s, **k: _swingup_cost(s).squeeze(), device="cpu", **kw)
    return model, ctrl


def test_multidisco_control_regulariser_on_cpu(monkeypatch):
    """Host side of ctrl_penalty != 1 (disco.py:90, 334-344): a_reg, the a_mat @ a_pre the kernel is handed, the plan
    update and `step("average")`, two rounds against the reference's recording."""
    d = load("ctrlpen_pendulum")
    S, N, H, A = d["actions0"].shape
    model, ctrl = _controller(monkeypatch, d, S, N, H, A, ctrl_penalty=float(d["ctrl_penalty"]), params_sampling=True,
                              params_samples=d["params0"].shape[0])
    ctrl.a_mat = d["a_mat0"].clone()
    for it in range(2):
        costs, states, actions, weights, _ = ctrl.forward(d["state"], model, _FixedParams(d[f"params{it}"]), d[f"actions{it}"])
        assert states.shape == (d["params0"].shape[0], S, N, H + 1, 2) and actions.shape == (d["params0"].shape[0], S, N, H, A)
        assert rel_max(costs, d[f"costs{it}"]) <= 1e-5
        assert rel_max(ctrl.a_mat, d[f"a_mat_fwd{it}"]) <= 1e-4 and rel_max(ctrl.a_mix, d[f"a_mix{it}"]) <= 1e-4
        assert rel_max(ctrl.step(strategy="average"), d[f"action{it}"]) <= 1e-4


@pytest.mark.parametrize("name", ["utf_pendulum_n1_gmm", "utf_pendulum_n3_mvn"])
def test_multidisco_sigma_points_on_cpu(monkeypatch, name):
    """Host side of params_sampling = MerweScaledUTF (disco.py:211-292): sigma points of the belief (both the
    `covariance_matrix` and the `variance.diag()` branch), their weighted log-probability, the reference's state row
    order and its un-tiled actions."""
    import torch.distributions as dist

    from dust_b200.utils.utf import MerweScaledUTF

    d = load(name)
    S, N, H, A = d["actions"].shape
    model, ctrl = _controller(monkeypatch, d, S, N, H, A, params_sampling=MerweScaledUTF(n=2, alpha=0.5), params_log_space=False)
    ctrl.a_mat = d["a_mat0"].clone()
    if bool(d["belief_is_mvn"]):
        pd = dist.MultivariateNormal(d["mean"], d["cov"])
    else:
        pd = dist.MixtureSameFamily(dist.Categorical(torch.ones(d["locs"].shape[0])),
                                    dist.Independent(dist.Normal(d["locs"], float(d["comp_sigma"])), 1))
    costs, states, actions, weights, plogp = ctrl.forward(d["state"], model, pd, d["actions"])
    assert tuple(actions.shape) == tuple(int(v) for v in d["acts_shape"])
    assert rel_max(states, d["states"]) <= 1e-6 and rel_max(costs, d["costs"]) <= 1e-6
    assert rel_max(plogp, d["params_log_p"]) <= 1e-5
    assert rel_max(ctrl.a_mat, d["a_mat1"]) <= 1e-4


def test_mpf_class_on_cpu(monkeypatch):
    """Host side of the parameter filter (mpf.py:12-86): conditioning on the observed transition, the prior whose
    centres alias the particles (H24), the bandwidth handed to the kernel -- the particle demo's filter, four
    control steps of 20 SVGD steps, with the oracle standing in for `dust_mpf_optimize`."""
    from dust_b200.inference import mpf as mpf_module
    from dust_b200.inference.likelihoods import GaussianLikelihood
    from dust_b200.models.particle import Particle
    from tests.util import golden_grid

    env = dict(dt=0.015, control_type="acceleration", noise_std=[0.1, 0.1], init_state=[-9.0, -9.0, 0, 0],
               target_state=[9.0, 9.0, 0, 0], can_crash=True, with_obstacle=True, deterministic=True,
               cost_params=dict(w_qpos=0.5, w_qvel=0.25, w_ctrl=0.2, w_obs=1.0e6, w_qpos_T=1.0e3, w_qvel_T=0.1),
               obst_preset="grid_4x4", obst_width=2.1, max_speed=5, max_accel=10, map_cell_size=0.1, map_size=[22, 22],
               map_type="direct")
    model_o = O.Model("particle", O.ParticleCfg(golden_grid()))

    class Ops:
        @staticmethod
        def mpf_optimize(spec, x, obs0, action, obs1, prior_inv_var, obs_std, bw, lr, n_steps, log_space, **kw):
            x1, gn = O.mpf_optimize(model_o, x[0].double(), obs0[0].double(), action[0].double(), obs1[0].double(), obs_std,
                                    (1.0 / prior_inv_var).double(), bw, lr, n_steps, log_space)
            x[0] = x1.float()                       # in place, as the kernel does
            return gn.float().unsqueeze(0)

    monkeypatch.setattr(mpf_module, "ops", Ops)
    d = load("dual_particle")
    model = Particle(**env, uncertain_params=["mass"], mass=2.0)
    lik = GaussianLikelihood(initial_obs=d["t0_in_state"], obs_std=float(d["obs_std"]), model=model, log_space=True)
    mpf = mpf_module.MPF(init_particles=d["t0_in_mpf_x0"].clone(), likelihood=lik, optimizer_class=torch.optim.SGD,
                         lr=float(d["mpf_lr"]), bw=float(d["mpf_prior_bw0"]), bw_scale=1.0, device="cpu")
    prior0 = mpf.prior                                 # captured once, as the demos do (particle_example.py:171)
    for t in range(int(d["n_steps"])):
        gn, bw = mpf.optimize(d[f"t{t}_out_a_seq"][0], d[f"t{t}_out_next_state"], bw=0.5, n_steps=20)
        assert rel_max(mpf.x, d[f"t{t}_out_mpf_x1"]) <= 5e-4 and gn.shape == (20,)
        assert rel_max(gn, d[f"t{t}_out_mpf_grad_norms"]) <= 5e-3
    assert torch.equal(prior0.component_distribution.base_dist.loc, mpf.x)
    # any other optimiser is built once over the particle tensor (mpf.py:23) and stepped on phi from the kernel
    adam = mpf_module.MPF(init_particles=d["t0_in_mpf_x0"].clone(), likelihood=lik, optimizer_class=torch.optim.Adam, device="cpu")
    assert isinstance(adam.optimizer, torch.optim.Adam) and adam.optimizer.param_groups[0]["params"][0] is adam.x
