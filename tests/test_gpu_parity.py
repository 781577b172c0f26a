"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
generated from the unmodified reference.  Tolerances are BASELINE.json's: costs / rollouts
<= 1e-5 relative (float32), phi and updated particles <= 1e-4 relative, indices exact."""
import math

import numpy as np
import pytest
import torch

from oracle import dust_oracle as O
from tests.util import RTOL_COST, RTOL_PHI, assert_close_to_reference, golden_grid, load, rel_elem, rel_max

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def env():
    from dust_b200 import _lib as L
    from dust_b200.models.particle import Particle
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    env_params = dict(dt=0.015, control_type="acceleration", noise_std=[0.1, 0.1], init_state=[-9.0, -9.0, 0, 0],
                      target_state=[9.0, 9.0, 0, 0], can_crash=True, with_obstacle=True, deterministic=True,
                      cost_params=dict(w_qpos=0.5, w_qvel=0.25, w_ctrl=0.2, w_obs=1.0e6, w_qpos_T=1.0e3, w_qvel_T=0.1),
                      obst_preset="grid_4x4", obst_width=2.1, max_speed=5, max_accel=10, map_cell_size=0.1,
                      map_size=[22, 22], map_type="direct")
    part = Particle(**env_params, uncertain_params=["mass"], mass=2.0)
    pend = PendulumModel(uncertain_params=("length", "mass"))
    return dict(
        L=L, part=part, pend=pend,
        spec=dict(particle=part.device_spec(part.default_inst_cost, part.default_term_cost, DEV),
                  pendulum=pend.device_spec(inst_cost, term_cost, DEV)),
        cfg=O.ParticleCfg(golden_grid()),
    )


def cu(t):
    return None if t is None else torch.as_tensor(t, dtype=torch.float32).to(DEV).contiguous()


def dev_params(kind, params, log_space):
    """sampled params (golden layout) -> (kernel params [1,P,dp], tiling)."""
    if params is None or params.numel() == 0:
        return None, 0
    p = params.exp() if log_space else params
    tiling = 1 if p.ndim == 1 else 0
    p = p.reshape(p.shape[0], -1)
    return cu(p).unsqueeze(0), tiling


# ---------------------------------------------------------------------------------------------
# K1: rollout + cost + soft-min
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["pendulum", "particle"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rollout_cost_vs_reference(env, kind, seed):
    from dust_b200 import ops

    d = load(f"fwd_{kind}_s{seed}")
    params, tiling = dev_params(kind, d["params"], bool(d["log_space"]))
    out = ops.rollout_cost(env["spec"][kind], cu(d["state"]).reshape(1, -1), cu(d["actions"]).unsqueeze(0),
                           params=params, param_tiling=tiling, temperature=float(d["temp"]),
                           want=("costs", "mppi_weights", "mppi_delta", "mix", "states"))
    costs = out["costs"][0].cpu()
    assert rel_elem(costs, d["costs"]) <= RTOL_COST
    states = out["states"][0].cpu()
    if kind == "particle":
        # +, *, /, floor, clamp only: trajectories are bit-identical to the CPU reference
        assert torch.equal(states[:, :8], d["states_sub"])
        assert torch.equal(states[..., -1, :], d["states_last"])
    else:
        assert rel_max(states[:, :8], d["states_sub"]) <= RTOL_COST
        assert rel_max(states[..., -1, :], d["states_last"]) <= RTOL_COST
    assert float((out["mppi_weights"][0].cpu() - d["weights"]).abs().max()) <= 2e-5
    assert rel_max(d["a_mat0"] + out["mppi_delta"][0].cpu(), d["a_mat1"]) <= 1e-4
    assert float((out["mix"][0].cpu() - d["a_mix1"]).abs().max()) <= 2e-5


def test_rollout_default_params_and_internal_sampling(env):
    from dust_b200 import ops

    d = load("fwd_pendulum_nops")  # N=8, S=256, H=20, P=1 (the batched-pendulum shape)
    out = ops.rollout_cost(env["spec"]["pendulum"], cu(d["state"]).reshape(1, -1), cu(d["actions"]).unsqueeze(0),
                           want=("costs",))
    assert rel_elem(out["costs"][0].cpu(), d["costs"]) <= RTOL_COST
    d = load("fwd_pendulum_internal")
    eps = d["eps"] * d["sigma"]
    out = ops.rollout_cost(env["spec"]["pendulum"], cu(d["state"]).reshape(1, -1), cu(eps + d["a_mat0"]).unsqueeze(0),
                           params=cu(d["params"]).unsqueeze(0), pert=cu(eps).unsqueeze(0),
                           want=("costs", "mppi_delta", "mix"))
    assert rel_elem(out["costs"][0].cpu(), d["costs"]) <= RTOL_COST
    assert rel_max(d["a_mat0"] + out["mppi_delta"][0].cpu(), d["a_mat1"]) <= 1e-4


@pytest.mark.parametrize("name,kind,log", [("svmpc_pendulum_rbf", "pendulum", False), ("svmpc_particle_rbf", "particle", True),
                                           ("dual_particle", "particle", True)])
def test_fused_actions_likelihood_and_analytic_gradient(env, name, kind, log):
    """theta + sigma*eps formed in-kernel; log-likelihood; svmpc.py:46-54 gradient."""
    from dust_b200 import ops

    d = load(name)
    for t in range(int(d["n_steps"])):
        gi, go = (lambda k: d[f"t{t}_in_{k}"]), (lambda k: d[f"t{t}_out_{k}"])
        params, tiling = dev_params(kind, gi("params"), log)
        out = ops.rollout_cost(env["spec"][kind], cu(gi("state")).reshape(1, -1), cu(gi("eps")).unsqueeze(0),
                               theta=cu(gi("theta0")).unsqueeze(0), sigma=cu(d["sigma"]), params=params,
                               param_tiling=tiling, alpha=1.0, want=("costs", "log_lik", "lik_weights", "grad_lik"))
        costs = out["costs"][0].cpu()
        assert rel_elem(costs, go("costs")) <= RTOL_COST
        assert rel_max(out["log_lik"][0].cpu(), go("log_l")) <= RTOL_COST
        actions = gi("theta0") + d["sigma"] * gi("eps")
        g_ref = O.analytic_lik_grad(go("costs").double(), actions.double(), gi("theta0").double(), d["sigma"].double(), 1.0)
        # the soft-min weights amplify cost differences by alpha: compare on the device's own costs too
        g_dev = O.analytic_lik_grad(costs.double(), actions.double(), gi("theta0").double(), d["sigma"].double(), 1.0)
        assert rel_max(out["grad_lik"][0].cpu(), g_dev) <= 1e-5
        assert rel_max(out["grad_lik"][0].cpu(), g_ref) <= 5e-3


def test_expected_cost_likelihood(env):
    from dust_b200 import ops

    d = load("fwd_pendulum_s0")
    L = env["L"]
    out = ops.rollout_cost(env["spec"]["pendulum"], cu(d["state"]).reshape(1, -1), cu(d["actions"]).unsqueeze(0),
                           params=cu(d["params"]).unsqueeze(0), likelihood=L.LIK_EXPECTED_COST, alpha=0.5,
                           want=("costs", "log_lik"))
    assert rel_max(out["log_lik"][0].cpu(), O.expected_cost_log_prob(d["costs"], 0.5)) <= RTOL_COST


def test_batched_instances_match_single_instance(env):
    """B instances in one launch == B separate launches (bitwise), ragged S*N tile edge included."""
    from dust_b200 import ops

    torch.manual_seed(0)
    B, S, N, H = 5, 37, 3, 11
    for kind, ds, A, dp in (("pendulum", 2, 1, 2), ("particle", 4, 2, 1)):
        state = torch.randn(B, ds, device=DEV)
        if kind == "particle":
            state = state * torch.tensor([6.0, 6.0, 1.0, 1.0], device=DEV)
        theta = torch.randn(B, N, H, A, device=DEV)
        eps = torch.randn(B, S, N, H, A, device=DEV)
        sigma = torch.full((A,), 2.0, device=DEV)
        params = torch.rand(B, 4, dp, device=DEV) + 0.7
        want = ("costs", "log_lik", "grad_lik", "mix")
        full = ops.rollout_cost(env["spec"][kind], state, eps, theta=theta, sigma=sigma, params=params, want=want)
        for b in range(B):
            one = ops.rollout_cost(env["spec"][kind], state[b:b + 1].contiguous(), eps[b:b + 1].contiguous(),
                                   theta=theta[b:b + 1].contiguous(), sigma=sigma, params=params[b:b + 1].contiguous(),
                                   want=want)
            for k in want:
                assert torch.equal(full[k][b], one[k][0]), (kind, k, b)
        # and against the oracle
        model = O.Model(kind, env["cfg"])
        for b in (0, B - 1):
            acts = (theta[b] + sigma * eps[b]).cpu()
            ref = O.disco_forward(model, state[b].cpu(), acts, params[b].cpu())
            assert rel_elem(full["costs"][b].cpu(), ref["costs"]) <= RTOL_COST


@pytest.mark.parametrize("kind,N,S,H,P", [("pendulum", 8, 67, 20, 0), ("pendulum", 3, 50, 30, 2), ("pendulum", 5, 9, 7, 3),
                                          ("particle", 6, 30, 16, 2), ("particle", 4, 70, 5, 0)])
def test_fused_instance_kernel_matches_two_stage_path_and_oracle(env, kind, N, S, H, P):
    """B >= 74 takes the fused one-launch kernel (online soft-min); a single instance takes the
    two-stage path.  Same costs bit for bit; likelihood and gradient to rounding."""
    from dust_b200 import ops

    torch.manual_seed(N * 1000 + S)
    B = 96
    ds, A, dp = (2, 1, 2) if kind == "pendulum" else (4, 2, 1)
    state = torch.randn(B, ds) * (torch.tensor([6.0, 6.0, 1.0, 1.0]) if kind == "particle" else torch.tensor([2.0, 2.0]))
    theta, eps = torch.randn(B, N, H, A) * 2, torch.randn(B, S, N, H, A)
    sigma = torch.full((A,), 1.7)
    params = cu(torch.rand(B, P, dp) + 0.7) if P else None
    L = env["L"]
    for lik, alpha in ((L.LIK_EXP_UTILITY, 0.9), (L.LIK_EXPECTED_COST, 0.3)):
        want = ("costs", "log_lik", "grad_lik")
        fused = ops.rollout_cost(env["spec"][kind], cu(state), cu(eps), theta=cu(theta), sigma=cu(sigma), params=params,
                                 likelihood=lik, alpha=alpha, want=want)
        for b in (0, 17, B - 1):
            one = ops.rollout_cost(env["spec"][kind], cu(state[b:b + 1]), cu(eps[b:b + 1]), theta=cu(theta[b:b + 1]),
                                   sigma=cu(sigma), params=None if params is None else params[b:b + 1].contiguous(),
                                   likelihood=lik, alpha=alpha, want=want)
            assert torch.equal(fused["costs"][b], one["costs"][0])
            assert rel_max(fused["log_lik"][b].cpu(), one["log_lik"][0].cpu()) <= 1e-6
            assert rel_max(fused["grad_lik"][b].cpu(), one["grad_lik"][0].cpu()) <= 1e-5
            costs = fused["costs"][b].cpu()
            acts = theta[b] + sigma * eps[b]
            ref = O.disco_forward(O.Model(kind, env["cfg"]), state[b], acts, None if params is None else params[b].cpu())
            assert rel_elem(costs, ref["costs"]) <= RTOL_COST
            g_dev = O.analytic_lik_grad(costs.double(), acts.double(), theta[b].double(), sigma.double(), alpha)
            assert rel_max(fused["grad_lik"][b].cpu(), g_dev) <= 1e-5


@pytest.mark.parametrize("N,S,H,P,tiling", [(4, 70, 18, 3, 0), (16, 33, 20, 2, 0), (32, 9, 12, 5, 1), (8, 64, 20, 4, 1),
                                            (64, 5, 6, 0, 0), (2, 200, 31, 2, 0)])
def test_packed_pendulum_kernel_shapes(env, N, S, H, P, tiling):
    """The two-trajectories-per-thread kernel (FFMA2/FADD2/FMUL2) across its branches: horizons that
    are not a multiple of 4, ragged last tiles (scalar fallback for an unpaired row), blocked and
    interleaved parameter tiling (the two rows of a thread then use different draws).  Every lane is
    IEEE-rounded like the scalar instruction: costs must equal the scalar staged path BIT FOR BIT."""
    from dust_b200 import ops

    torch.manual_seed(N * 131 + S)
    B = 80
    state = torch.randn(B, 2) * torch.tensor([2.5, 2.0])
    theta, eps = torch.randn(B, N, H, 1) * 2, torch.randn(B, S, N, H, 1)
    sigma = torch.tensor([1.3])
    params = cu(torch.rand(B, P, 2) * 0.7 + 0.6) if P else None
    kw = dict(theta=cu(theta), sigma=cu(sigma), params=params, param_tiling=tiling, alpha=0.8,
              want=("costs", "log_lik", "grad_lik"))
    packed = ops.rollout_cost(env["spec"]["pendulum"], cu(state), cu(eps), **kw)
    for b in (0, B // 2, B - 1):
        one = ops.rollout_cost(env["spec"]["pendulum"], cu(state[b:b + 1]), cu(eps[b:b + 1]), theta=cu(theta[b:b + 1]),
                               sigma=cu(sigma), params=None if params is None else params[b:b + 1].contiguous(),
                               param_tiling=tiling, alpha=0.8, want=("costs", "log_lik", "grad_lik"))
        assert torch.equal(packed["costs"][b], one["costs"][0]), (N, S, H, P, tiling, b)
        assert rel_max(packed["log_lik"][b].cpu(), one["log_lik"][0].cpu()) <= 1e-6
        assert rel_max(packed["grad_lik"][b].cpu(), one["grad_lik"][0].cpu()) <= 1e-5


# ---------------------------------------------------------------------------------------------
# K3 GMM prior, K5 phi, K7 forward
# ---------------------------------------------------------------------------------------------
def test_gmm_score_and_log_prob(env):
    from dust_b200 import ops

    torch.manual_seed(1)
    for (M, K, D) in ((3, 3, 30), (6, 6, 80), (50, 50, 2), (8, 8, 20), (33, 17, 200)):
        x, mu = torch.randn(2, M, D), torch.randn(2, K, D)
        mix = torch.rand(2, K)
        mix[1, 0] = 0.0  # zero weight -> clamped to eps, not -inf (torch Categorical semantics)
        var = torch.rand(D) + 0.5
        lp, sc = ops.gmm(cu(x), cu(mu), cu(mix), cu(1.0 / var), ops.gmm_log_norm(var))
        for b in range(2):
            assert rel_max(lp[b].cpu(), O.gmm_log_prob(x[b].double(), mu[b].double(), mix[b].double(), var.double())) <= 1e-5
            assert rel_max(sc[b].cpu(), O.gmm_score(x[b].double(), mu[b].double(), mix[b].double(), var.double())) <= 1e-4
        lp2, _ = ops.gmm(cu(x), cu(mu), None, cu(1.0 / var), ops.gmm_log_norm(var), want_score=False)
        assert rel_max(lp2[0].cpu(), O.gmm_log_prob(x[0].double(), mu[0].double(), torch.ones(K).double(), var.double())) <= 1e-5


@pytest.mark.parametrize("N,D", [(3, 30), (6, 80), (8, 20), (64, 40), (200, 7)])
def test_phi_small_all_variants(env, N, D):
    from dust_b200 import ops

    torch.manual_seed(N * 100 + D)
    x, s = torch.randn(2, N, D) * 2.0, torch.randn(2, N, D)
    for (gamma, c1, c2) in ((1 / (2 * O.GPYTORCH_DEFAULT_LENGTHSCALE ** 2), 1 / N, -1 / O.GPYTORCH_DEFAULT_LENGTHSCALE ** 2),
                            (1 / (2 * 1.5 ** 2), 1 / N, 1 / (N * 1.5 ** 2)), (0.01, 1 / N, 2 * 0.01 / N)):
        out = ops.svgd_phi(cu(x), cu(s), gamma=gamma, c1=c1, c2=c2, lr=0.25, want_update=True)
        for b in range(2):
            ref = O.phi_unified(x[b].double(), s[b].double(), gamma, c1, c2)
            assert rel_max(out["phi"][b].cpu(), ref) <= RTOL_PHI
            assert rel_max(out["x_out"][b].cpu(), x[b].double() + 0.25 * ref) <= RTOL_PHI
    if N <= 64:
        out = ops.svgd_phi(cu(x), cu(s), per_dim=True, want_bandwidths=True)
        for b in range(2):
            _, _, hs = O.iid_mp_eval(x[b], x[b].clone())
            assert rel_max(out["bandwidths"][b].cpu(), hs) <= 1e-6
            assert rel_max(out["phi"][b].cpu(), O.phi_svmpc_iid_mp(x[b], s[b])) <= RTOL_PHI


def test_phi_row_range(env):
    from dust_b200 import ops

    torch.manual_seed(2)
    x, s = cu(torch.randn(1, 40, 12)), cu(torch.randn(1, 40, 12))
    full = ops.svgd_phi(x, s, gamma=0.05, c1=1 / 40, c2=0.01)["phi"]
    part = torch.zeros_like(full)
    for r in ((0, 13), (13, 40)):
        a = ops.svgd_phi(x, s, gamma=0.05, c1=1 / 40, c2=0.01, rows=r)["phi"]
        part[:, r[0]:r[1]] = a[:, r[0]:r[1]]
    assert torch.equal(full, part)


SEQ = [("svmpc_pendulum_rbf", "pendulum", "gpytorch"), ("svmpc_pendulum_mp", "pendulum", "mp"),
       ("svmpc_particle_rbf", "particle", "gpytorch"), ("svmpc_particle_mp", "particle", "mp"),
       ("dual_pendulum_bw", "pendulum", "gpytorch"), ("dual_particle", "particle", "gpytorch")]
HYPER = {"pendulum": dict(alpha=1.0, lr=2.0, var=4.0, wp=False, log=False),
         "particle": dict(alpha=1.0, lr=100.0, var=25.0, wp=True, log=True)}


@pytest.mark.parametrize("name,kind,kern", SEQ)
def test_svmpc_control_step_teacher_forced(env, name, kind, kern):
    """SVMPC.optimize + SVMPC.forward on the recorded inputs of every control step."""
    from dust_b200.inference.core import SvmpcCore

    d = load(name)
    c = HYPER[kind]
    model = O.Model(kind, env["cfg"])
    for t in range(int(d["n_steps"])):
        gi, go = (lambda k: d[f"t{t}_in_{k}"]), (lambda k: d[f"t{t}_out_{k}"])
        params, tiling = dev_params(kind, gi("params"), c["log"])
        A = gi("theta0").shape[-1]
        core = SvmpcCore(env["spec"][kind], cu(gi("theta0")).unsqueeze(0), cu(gi("mu0")).unsqueeze(0),
                         cu(gi("mix0")).unsqueeze(0), torch.full((A,), c["var"]), d["sigma"], alpha=c["alpha"],
                         temperature=1.0 / c["alpha"], lr=c["lr"], kernel=kern, weighted_prior=c["wp"], aliased=t > 0)
        out = core.optimize_step(cu(gi("state")).reshape(1, -1), cu(gi("eps")).unsqueeze(0), params, tiling)
        theta1 = core.theta[0].cpu()
        a_seq, pw, i_star = core.forward_step()
        # float64 restatement on the same inputs
        st = O.SvmpcState(gi("theta0").double(), gi("mu0").double(), gi("mix0").double(), c["var"], aliased=t > 0)
        p64 = gi("params").double() if gi("params").numel() else None
        ref = O.svmpc_optimize(model, st, gi("state").double(), gi("eps").double(), d["sigma"].double(), p64,
                               c["log"], c["alpha"], c["lr"], kernel="rbf" if kern == "gpytorch" else "mp")
        assert rel_elem(out["costs"][0].cpu(), go("costs")) <= RTOL_COST
        assert_close_to_reference(out["phi"][0].cpu(), go("phi"), ref["phi"], RTOL_PHI, f"{name} t{t} phi")
        assert_close_to_reference(theta1, go("theta1"), ref["theta1"], RTOL_PHI, f"{name} t{t} theta1")
        assert int(i_star[0]) == int(go("i_star"))
        assert float((pw[0].cpu() - go("p_weights")).abs().max()) <= 1e-4
        assert torch.equal(a_seq[0].cpu(), theta1[int(go("i_star"))])
        rolled = theta1.roll(-1, dims=-2)
        rolled[..., -1, :] = rolled[..., -2, :]
        assert torch.equal(core.theta[0].cpu(), rolled)
        mix2 = core.mix[0].cpu()
        assert float((mix2 / mix2.sum() - go("mix2")).abs().max()) <= 1e-4


def test_forward_mean_roll_and_argmax_ties(env):
    from dust_b200 import ops

    B, N, H, A = 3, 5, 7, 2
    torch.manual_seed(3)
    theta = torch.randn(B, N, H, A)
    ll = torch.zeros(B, N)  # identical particles -> ties -> first index wins
    theta[1] = theta[1, 0]
    out = ops.svmpc_forward(cu(ll), cu(theta), cu(theta), None, cu(torch.ones(H * A)), 0.0, roll_strategy=1)
    assert int(out["i_star"][1]) == 0
    exp = theta.roll(-1, dims=-2)
    exp[..., -1, :] = theta.mean(dim=-2)
    assert rel_max(out["theta_next"].cpu(), exp) <= 1e-6
    assert torch.equal(out["mix_next"].cpu(), torch.ones(B, N))


@pytest.mark.parametrize("kind", ["pendulum", "particle"])
def test_disco_step(env, kind):
    from dust_b200 import ops

    d = load(f"fwd_{kind}_s0")
    lim = 2.0 if kind == "pendulum" else 10.0
    A = d["a_mat1"].shape[-1]
    lo, hi = cu(torch.full((A,), -lim)), cu(torch.full((A,), lim))
    for strat, code in (("argmax", 0), ("average", 1)):
        a_mat = cu(d["a_mat1"]).unsqueeze(0).clone()
        nxt, a_seq = ops.disco_step(a_mat, cu(d["a_mix1"]).unsqueeze(0), lo, hi, code, 1)
        assert rel_max(nxt[0].cpu(), d[f"step_{strat}_action"]) <= 1e-6
        assert rel_max(a_seq[0].cpu(), d[f"step_{strat}_a_seq"]) <= 1e-6
        assert rel_max(a_mat[0].cpu(), d[f"step_{strat}_a_mat"]) <= 1e-6


# ---------------------------------------------------------------------------------------------
# K2 adjoint
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kind,log", [("pathwise_pendulum_s4", "pendulum", False), ("pathwise_particle_s4", "particle", True),
                                           ("pathwise_particle_s5", "particle", True)])
def test_adjoint_matches_autograd_of_reference(env, name, kind, log):
    from dust_b200 import ops

    d = load(name)
    params, tiling = dev_params(kind, d["params"], log)
    args = dict(theta=cu(d["theta"]).unsqueeze(0), sigma=cu(d["sigma"]), params=params, param_tiling=tiling)
    fwd = ops.rollout_cost(env["spec"][kind], cu(d["state"]).reshape(1, -1), cu(d["eps"]).unsqueeze(0), alpha=1.0,
                           want=("costs", "log_lik", "lik_weights"), **args)
    g = ops.rollout_adjoint(env["spec"][kind], cu(d["state"]).reshape(1, -1), cu(d["eps"]).unsqueeze(0),
                            fwd["lik_weights"], alpha=1.0, **args)
    assert rel_elem(fwd["costs"][0].cpu(), d["costs"]) <= RTOL_COST
    g64 = O.pathwise_lik_grad_adjoint(O.Model(kind, env["cfg"]), d["state"].double(), d["theta"].double(), d["eps"].double(),
                                      d["sigma"].double(), d["params"].double(), log, 1.0)[0]
    assert_close_to_reference(g[0].cpu(), d["grad"], g64, RTOL_PHI, name)


def test_adjoint_batched_and_expected_cost(env):
    from dust_b200 import ops

    torch.manual_seed(5)
    L = env["L"]
    B, S, N, H = 3, 19, 4, 9
    for kind, ds, A, dp in (("pendulum", 2, 1, 2), ("particle", 4, 2, 1)):
        state = torch.randn(B, ds) * (torch.tensor([6.0, 6.0, 1.0, 1.0]) if kind == "particle" else 1.0)
        theta, eps = torch.randn(B, N, H, A), torch.randn(B, S, N, H, A)
        sigma = torch.full((A,), 1.5)
        params = torch.rand(B, 3, dp) + 0.7
        model = O.Model(kind, env["cfg"])
        for lik in (L.LIK_EXP_UTILITY, L.LIK_EXPECTED_COST):
            fwd = ops.rollout_cost(env["spec"][kind], cu(state), cu(eps), theta=cu(theta), sigma=cu(sigma), params=cu(params),
                                   likelihood=lik, alpha=0.7, want=("costs", "lik_weights"))
            g = ops.rollout_adjoint(env["spec"][kind], cu(state), cu(eps), fwd["lik_weights"], theta=cu(theta),
                                    sigma=cu(sigma), params=cu(params), likelihood=lik, alpha=0.7)
            w_dev = fwd["lik_weights"].cpu().double()
            for b in range(B):
                # the adjoint is linear in the soft-min weights: hold the device's weights fixed so the
                # comparison is not dominated by float32 rounding of costs ~1e5 inside exp(-alpha C)
                x = theta[b].double().clone().requires_grad_(True)
                out = O.disco_forward(model, state[b].double(), x + sigma.double() * eps[b].double(), params[b].double())
                if lik == L.LIK_EXP_UTILITY:
                    obj = (-0.7 * w_dev[b] * out["costs"]).sum()
                else:
                    obj = O.expected_cost_log_prob(out["costs"], 0.7).sum()
                (gr,) = torch.autograd.grad(obj, x)
                assert rel_max(g[b].cpu(), gr) <= RTOL_PHI, (kind, lik, b)


@pytest.mark.parametrize("kind,ds,A,dp", [("pendulum", 2, 1, 2), ("particle", 4, 2, 1)])
def test_adjoint_with_sparse_weights(env, kind, ds, A, dp):
    """Rows whose soft-min weight is exactly zero are not rolled out, dead tiles return early, and
    sparse tiles spread their live rows over the idle threads: the gradient must equal the autograd
    of the oracle under the SAME weights.  Weight patterns: one live row in the whole problem, a few
    per tile, a dense tile next to dead ones, everything live."""
    from dust_b200 import ops

    torch.manual_seed(17)
    B, N, S, H, P = 2, 4, 97, 9, 5                      # S*N = 388 rows: three full tiles + a ragged one
    state = torch.randn(B, ds) * (torch.tensor([6.0, 6.0, 1.0, 1.0]) if kind == "particle" else 1.0)
    theta, eps = torch.randn(B, N, H, A), torch.randn(B, S, N, H, A)
    sigma, params = torch.full((A,), 1.5), torch.rand(B, P, dp) + 0.7
    model = O.Model(kind, env["cfg"])
    patterns = {}
    w = torch.zeros(B, S, N); w[0, 40, 2] = 1.0; w[1, 96, 3] = 0.25
    patterns["one_row"] = w
    w = torch.zeros(B, S, N); idx = torch.randperm(S * N)[:23]; w.view(B, -1)[:, idx] = torch.rand(B, 23)
    patterns["few_per_tile"] = w
    w = torch.zeros(B, S, N); w.view(B, -1)[:, 128:256] = torch.rand(B, 128); w.view(B, -1)[:, 300] = 0.5
    patterns["dense_tile_among_dead"] = w
    patterns["all_live"] = torch.rand(B, S, N) + 0.01
    for name, w in patterns.items():
        g = ops.rollout_adjoint(env["spec"][kind], cu(state), cu(eps), cu(w), theta=cu(theta), sigma=cu(sigma), params=cu(params),
                                alpha=0.7)
        for b in range(B):
            x = theta[b].double().clone().requires_grad_(True)
            out = O.disco_forward(model, state[b].double(), x + sigma.double() * eps[b].double(), params[b].double())
            (gr,) = torch.autograd.grad((-0.7 * w[b].double() * out["costs"]).sum(), x)
            assert rel_max(g[b].cpu(), gr) <= RTOL_PHI, (kind, name, b)
    # no live row at all: the gradient is exactly zero
    g0 = ops.rollout_adjoint(env["spec"][kind], cu(state), cu(eps), cu(torch.zeros(B, S, N)), theta=cu(theta), sigma=cu(sigma),
                             params=cu(params), alpha=0.7)
    assert float(g0.abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------
# MPF
# ---------------------------------------------------------------------------------------------
def test_mpf_against_reference(env):
    from dust_b200 import ops

    d = load("mpf_pendulum")
    x = cu(d["x0"]).unsqueeze(0).clone()
    gn = ops.mpf_optimize(env["spec"]["pendulum"], x, cu(d["obs0"]).reshape(1, -1), cu(d["action"]).reshape(1, -1),
                          cu(d["obs1"]).reshape(1, -1), cu(torch.full((2,), 1 / 0.01)), 0.1, 0.1, 1e-3, 1, False)
    phi0 = (x[0].cpu() - d["x0"]) / 1e-3
    assert rel_max(phi0, d["phi0"]) <= 2e-3  # difference quotient of a float32 update
    assert rel_max(gn[0, 0].cpu(), d["phi0"].norm()) <= RTOL_PHI
    # 20 steps: unstable iteration (amplifies rounding ~1.4x per step) -> conditioning-aware bound
    x = cu(d["x0"]).unsqueeze(0).clone()
    ops.mpf_optimize(env["spec"]["pendulum"], x, cu(d["obs0"]).reshape(1, -1), cu(d["action"]).reshape(1, -1),
                     cu(d["obs1"]).reshape(1, -1), cu(torch.full((2,), 1 / 0.01)), 0.1, 0.1, 1e-3, 20, False)
    dd = {k: v.double() for k, v in d.items()}
    x64, _ = O.mpf_optimize(O.Model("pendulum"), dd["x0"], dd["obs0"], dd["action"], dd["obs1"], 0.1, 0.01, 0.1, 1e-3, 20, False)
    assert_close_to_reference(x[0].cpu(), d["x1"], x64, RTOL_PHI, "mpf pendulum x1")
    d = load("mpf_particle")
    x = cu(d["x0"]).unsqueeze(0).clone()
    gn = ops.mpf_optimize(env["spec"]["particle"], x, cu(d["obs0"]).reshape(1, -1), cu(d["action"]).reshape(1, -1),
                          cu(d["obs1"]).reshape(1, -1), cu(torch.full((1,), 1 / 0.01)), 0.1, 0.5, 0.01, 20, True)
    assert rel_max(x[0].cpu(), d["x1"]) <= RTOL_PHI
    assert rel_max(gn[0].cpu(), d["grad_norms"]) <= RTOL_PHI


@pytest.mark.parametrize("kind,Np,steps", [("particle", 512, 20), ("pendulum", 300, 7), ("particle", 131, 1)])
def test_mpf_cooperative_kernel_equals_single_cta(env, kind, Np, steps):
    """One large instance spread over many SMs (cooperative launch, two grid barriers per step): every
    particle keeps its own warp and lane order, so the particles come out BIT-IDENTICAL to the
    one-CTA kernel; odd step counts leave the result in the workspace and copy it back."""
    from dust_b200 import ops

    torch.manual_seed(Np)
    spec = env["spec"][kind]
    if kind == "particle":
        x0 = math.log(2.0) + 0.1 * torch.randn(1, Np, 1)
        obs0, act = torch.tensor([[-9.0, -9.0, 0.3, 0.1]]), torch.tensor([[4.0, -3.0]])
        log_space, lr, bw = True, 1e-2, 0.5
        obs1 = ops.model_step(spec, cu(obs0), cu(act), cu(torch.tensor([[2.6]]))).cpu()
    else:
        x0 = 0.6 + 0.7 * torch.rand(1, Np, 2)
        obs0, act = torch.tensor([[3.0, 0.2]]), torch.tensor([[1.5]])
        log_space, lr, bw = False, 1e-3, 0.05
        obs1 = ops.model_step(spec, cu(obs0), cu(act), cu(torch.tensor([[1.1, 0.8]]))).cpu()
    piv = cu(torch.full((spec.dp,), 1.0 / bw ** 2))
    outs = []
    for coop in (True, False):
        x = cu(x0).clone()
        gn = ops.mpf_optimize(spec, x, cu(obs0), cu(act), cu(obs1), piv, 0.1, bw, lr, steps, log_space, cooperative=coop)
        outs.append((x.cpu(), gn.cpu()))
    assert not torch.equal(outs[0][0], x0)
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel_max(outs[0][1], outs[1][1]) <= 1e-5


def test_mpf_dual_loop_steps(env):
    from dust_b200 import ops

    d = load("dual_particle")
    for t in range(int(d["n_steps"])):
        pv = float(d["mpf_prior_bw0"]) ** 2 if t == 0 else float(d[f"t{t-1}_out_mpf_bw"]) ** 2
        x = cu(d[f"t{t}_in_mpf_x0"]).unsqueeze(0).clone()
        ops.mpf_optimize(env["spec"]["particle"], x, cu(d[f"t{t}_in_state"]).reshape(1, -1),
                         cu(d[f"t{t}_out_a_seq"][0]).reshape(1, -1), cu(d[f"t{t}_out_next_state"]).reshape(1, -1),
                         cu(torch.full((1,), 1 / pv)), float(d["obs_std"]), float(d[f"t{t}_out_mpf_bw"]),
                         float(d["mpf_lr"]), 20, True)
        assert rel_max(x[0].cpu(), d[f"t{t}_out_mpf_x1"]) <= RTOL_PHI


# ---------------------------------------------------------------------------------------------
# K4 median + large-N phi
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [64, 257, 1024])
def test_exact_median_and_svgd_phi_vs_reference(env, N):
    from dust_b200 import ops

    d = load(f"phi_svgd_N{N}")
    x = cu(d["X"])
    med = float(ops.median_sq_dist(x)[0])
    # exact rank selection over the device's own float32 distances; the value agrees with the
    # reference's MKL distances to rounding
    assert abs(med - float(d["median_d2"])) <= 4e-6 * float(d["median_d2"])
    bw = float(d["bw"])
    out = ops.svgd_phi(x.unsqueeze(0), cu(d["score"]).unsqueeze(0), gamma=1 / (2 * bw * bw), c1=1 / N, c2=1 / (N * bw * bw))
    p64 = O.phi_svgd(d["X"].double(), d["score"].double(), bw)
    assert_close_to_reference(out["phi"][0].cpu(), d["phi"], p64, RTOL_PHI, f"phi N={N}")


def test_median_is_exact_rank_statistic(env):
    """Against a sort of the same float32 distance formula (odd/even N, duplicates, ragged tiles)."""
    from dust_b200 import ops

    torch.manual_seed(7)
    for N, D in ((65, 3), (130, 40), (513, 8), (700, 17)):
        X = torch.randn(N, D) * 3
        X[N // 2:N // 2 + 5] = X[0]  # exact duplicates -> zero distances off the diagonal
        med = float(ops.median_sq_dist(cu(X))[0])
        d2 = O.sq_dists_addmm(X.double(), X.double())
        srt = d2.reshape(-1).sort().values
        k = (N * N - 1) // 2
        lo, hi = float(srt[max(k - 3, 0)]), float(srt[min(k + 3, N * N - 1)])
        assert lo - 1e-4 * abs(lo) - 1e-5 <= med <= hi + 1e-4 * abs(hi) + 1e-5


def _median_cases():
    g = torch.Generator().manual_seed(23)
    rn = lambda *s: torch.randn(*s, generator=g)
    cauchy = torch.tan(math.pi * (torch.rand(4096, 16, generator=g) - 0.5))
    return {
        "normal_2048": rn(2048, 40),
        "normal_8192": rn(8192, 40),
        "anisotropic": rn(4096, 40) * (torch.arange(1, 41).float() / 10),
        "low_dim": rn(4096, 8),
        "heavy_tailed": cauchy,
        "two_clusters": torch.cat([rn(2048, 24), rn(2048, 24) + 30.0]),
        "all_identical": torch.full((2048, 40), 0.3),
        "mostly_duplicates": torch.cat([torch.zeros(3072, 16), rn(1024, 16)]),
    }


@pytest.mark.parametrize("case", list(_median_cases()))
def test_tensor_core_median_vs_radix_select(env, case):
    """The sampled-window tensor-core pass and the two-pass radix select are both exact rank
    selections, over distances formed in a different order (3xTF32 Gram vs fp32 FMA chain): they
    agree to distance rounding.  Degenerate inputs (zero median, ties wider than the window) must
    either succeed or hand over on the device; either way the value returned is the radix one."""
    from dust_b200 import ops

    x = cu(_median_cases()[case])
    N, D = x.shape
    ws = ops.MedianWorkspace(N, D, x.device)
    assert ws.fast, "shape should qualify for the tensor-core pass"
    fast = float(ops.median_sq_dist(x, ws=ws)[0])
    took_fast = int(ws.selected[5])
    robust = float(ops.median_sq_dist(x, allow_fast=False)[0])
    assert abs(fast - robust) <= 4e-6 * abs(robust) + 1e-12, (case, fast, robust, took_fast)
    if case.startswith("normal") or case in ("anisotropic", "low_dim"):
        assert took_fast == 1, "well-behaved cloud should not need the fallback"
    # float64 sort of the same points brackets both
    X = _median_cases()[case].double()
    if N <= 4096:
        srt = O.sq_dists_addmm(X, X).reshape(-1).sort().values
        k = (N * N - 1) // 2
        lo, hi = float(srt[k - 8]), float(srt[k + 8])
        assert lo - 1e-5 * abs(lo) - 1e-6 <= fast <= hi + 1e-5 * abs(hi) + 1e-6


def test_large_phi_properties(env):
    """N = 4096: fused kernel vs float64 tiles; translation invariance; device-side bandwidth."""
    from dust_b200 import ops

    torch.manual_seed(11)
    N, D = 4096, 40
    X = torch.randn(N, D)
    S = -X
    x, s = cu(X).unsqueeze(0), cu(S).unsqueeze(0)
    med = ops.median_sq_dist(x[0])
    coef = ops.bandwidth_from_median(med, N, 1.0, 0)
    out = ops.svgd_phi(x, s, gamma_dev=coef)
    g, c1, c2, bw = [float(v) for v in coef.cpu()]
    ref = O.phi_unified_tiled(X, S, g, c1, c2, tile=1024)
    assert rel_max(out["phi"][0].cpu(), ref) <= RTOL_PHI
    bw_ref, med_ref = O.bw_median(X)
    assert abs(float(med[0]) - float(med_ref)) <= 4e-6 * float(med_ref)
    assert abs(bw - float(bw_ref)) <= 1e-5 * float(bw_ref)
    # rows split in blocks (what each rank computes) == the full result, to rounding: the number of
    # column splits per row tile depends on how many tiles a call covers, so partial sums group differently
    a = ops.svgd_phi(x, s, gamma_dev=coef, rows=(0, 1536))["phi"]
    b = ops.svgd_phi(x, s, gamma_dev=coef, rows=(1536, N))["phi"]
    assert rel_max(torch.cat([a[:, :1536], b[:, 1536:]], 1).cpu(), out["phi"].cpu()) <= 2e-6
    a = ops.svgd_phi(x, s, gamma_dev=coef, rows=(0, 1500))["phi"]
    b = ops.svgd_phi(x, s, gamma_dev=coef, rows=(1500, N))["phi"]
    assert rel_max(torch.cat([a[:, :1500], b[:, 1500:]], 1).cpu(), ref) <= RTOL_PHI


# ---------------------------------------------------------------------------------------------
# one instance sharded over its parameter draws (what each rank of ShardedRollout computes)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,P,world", [("pendulum", 8, 2), ("particle", 37, 3), ("particle", 5, 8)])
def test_parameter_draw_shares_sum_to_the_full_rollout(env, kind, P, world):
    """Draw ranges [p0, p1) (blocked and -- particle, scalar-event belief -- interleaved tiling) give their share of
    the mean cost and of the pathwise gradient; the shares of all ranks sum to the single-call result, and
    dust_cost_reduce on the summed costs reproduces its likelihood, weights and analytic gradient."""
    from dust_b200 import ops
    from dust_b200.distributed import row_block

    torch.manual_seed(P)
    spec = env["spec"][kind]
    S, N, H = 48, 4, 14
    if kind == "pendulum":
        state0, sigma, A = cu(torch.tensor([[2.5, 0.3]])), cu(torch.tensor([2.0])), 1
        params, tiling, alpha = cu(torch.rand(1, P, 2) * 0.7 + 0.6), 0, 0.05
    else:
        state0, sigma, A = cu(torch.tensor([[-6.0, -7.0, 0.5, 0.8]])), cu(torch.tensor([5.0, 5.0])), 2
        params, tiling, alpha = cu(torch.randn(1, P, 1) * 0.1 + math.log(2.0)).exp(), 1, 1e-4
    theta, noise = cu(torch.randn(1, N, H, A)), cu(torch.randn(1, S, N, H, A))
    kw = dict(theta=theta, sigma=sigma, params=params, param_tiling=tiling, alpha=alpha)
    full = ops.rollout_cost(spec, state0, noise, want=("costs", "log_lik", "lik_weights", "grad_lik"), **kw)
    g_full = ops.rollout_adjoint(spec, state0, noise, full["lik_weights"], **kw)
    costs = torch.zeros_like(full["costs"])
    grad = torch.zeros_like(g_full)
    for r in range(world):
        p0, p1 = row_block(P, r, world)
        if p1 == p0:
            continue
        costs += ops.rollout_cost(spec, state0, noise, want=("costs",), p_range=(p0, p1), **kw)["costs"]
        grad += ops.rollout_adjoint(spec, state0, noise, full["lik_weights"], p_range=(p0, p1), **kw)
    assert rel_elem(costs.cpu(), full["costs"].cpu()) <= 2e-6
    assert rel_max(grad.cpu(), g_full.cpu()) <= 1e-5
    red = ops.rollout_cost(spec, state0, noise, want=("log_lik", "lik_weights", "grad_lik"), out={"costs": full["costs"].clone()},
                           reduce_only=True, **kw)
    for k in ("log_lik", "lik_weights", "grad_lik"):
        assert torch.equal(red[k], full[k]), k
    with pytest.raises(ValueError):
        ops.rollout_cost(spec, state0, noise, want=("costs", "log_lik"), p_range=(0, 1), **kw)
    with pytest.raises(ValueError):
        ops.rollout_cost(spec, state0, noise, want=("costs",), p_range=(0, P + 1), **kw)


# ---------------------------------------------------------------------------------------------
# tcgen05 / TMEM 3xTF32 phi (N multiple of 128, N >= 1024, D <= 47)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,D,gamma", [(1024, 40, 0.02), (2048, 16, 0.05), (4096, 40, 0.015), (8192, 24, 0.03), (3072, 40, 1.5)])
def test_tensor_core_phi_vs_float64_and_simt(env, N, D, gamma):
    """wide kernels (every pair contributes) and a very narrow one (gamma = 1.5: only the exact-zero
    diagonal and near neighbours survive) against the tiled float64 oracle; SIMT tiles as a second opinion."""
    import os

    from dust_b200 import ops

    torch.manual_seed(N + D)
    X, S = torch.randn(N, D), torch.randn(N, D)
    x, s = cu(X).unsqueeze(0), cu(S).unsqueeze(0)
    c1, c2 = 1.0 / N, 0.37 / N
    os.environ.pop("DUST_B200_NO_TC", None)
    tc = ops.svgd_phi(x, s, gamma=gamma, c1=c1, c2=c2, lr=0.5, want_update=True)
    os.environ["DUST_B200_NO_TC"] = "1"
    try:
        simt = ops.svgd_phi(x, s, gamma=gamma, c1=c1, c2=c2)["phi"][0].cpu()
    finally:
        os.environ.pop("DUST_B200_NO_TC", None)
    ref = O.phi_unified_tiled(X, S, gamma, c1, c2, tile=1024)
    assert rel_max(tc["phi"][0].cpu(), ref) <= RTOL_PHI
    assert rel_max(simt, ref) <= RTOL_PHI
    assert rel_max(tc["x_out"][0].cpu(), X.double() + 0.5 * ref) <= RTOL_PHI
    # row blocks (what each rank computes) reproduce the full result to rounding
    a = ops.svgd_phi(x, s, gamma=gamma, c1=c1, c2=c2, rows=(0, 384))["phi"]
    b = ops.svgd_phi(x, s, gamma=gamma, c1=c1, c2=c2, rows=(384, N))["phi"]
    assert rel_max(torch.cat([a[:, :384], b[:, 384:]], 1).cpu(), tc["phi"].cpu()) <= 2e-6


def test_tensor_core_path_is_taken(env):
    """the profiler must show phi_tc_kernel (not the SIMT tiles) for a qualifying shape"""
    from dust_b200 import ops

    L = env["L"]
    lib = L.load()
    x = torch.randn(1, 2048, 40, device=DEV)
    lib.dust_profiler_reset()
    lib.dust_profiler_enable(1)
    ops.svgd_phi(x, -x, gamma=0.02, c1=1 / 2048, c2=1 / 2048)
    prof = L.profiler_report()
    lib.dust_profiler_enable(0)
    assert "phi_tc_kernel" in prof and "phi_large_kernel" not in prof


# ---------------------------------------------------------------------------------------------
# action-noise generator (Philox4x32-10 + Box-Muller)
# ---------------------------------------------------------------------------------------------
def _philox4x32_10(index, offset, seed):
    """numpy restatement of the published Philox4x32-10 (Salmon et al., SC'11) for known-answer checks."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [index & 0xFFFFFFFF, index >> 32, offset & 0xFFFFFFFF, offset >> 32]
    k = [seed & 0xFFFFFFFF, seed >> 32]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox_known_answers():
    """Random123's published vectors for philox4x32-10: counter/key all zero, and all ones."""
    assert _philox4x32_10(0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    ones = (1 << 64) - 1
    assert _philox4x32_10(ones, ones, ones) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_noise_generator_matches_philox_and_is_normal(env):
    from dust_b200 import ops

    # bit-level: the first blocks against the host restatement pushed through the same Box-Muller
    seed, offset = 0x1234_5678_9ABC_DEF0, 77
    out = ops.noise_normal(torch.empty(64, device=DEV), seed, offset).cpu().double()
    for i in range(16):
        w = _philox4x32_10(i, offset, seed)
        for h, (a, b) in enumerate(((w[0], w[1]), (w[2], w[3]))):
            u = (a + 1) * 2.0 ** -32
            rev = (b >> 9) * 2.0 ** -23
            r = math.sqrt(-2.0 * math.log(u))
            exp = (r * math.cos(2 * math.pi * rev), r * math.sin(2 * math.pi * rev))
            got = out[4 * i + 2 * h: 4 * i + 2 * h + 2]
            assert abs(float(got[0]) - exp[0]) <= 2e-5 * max(1.0, r) and abs(float(got[1]) - exp[1]) <= 2e-5 * max(1.0, r)
    # reproducible; other offsets / seeds give other streams; ragged length
    a = ops.noise_normal(torch.empty(1003, device=DEV), 5, 9)
    assert torch.equal(a, ops.noise_normal(torch.empty(1003, device=DEV), 5, 9))
    assert torch.equal(a[:1000], ops.noise_normal(torch.empty(1000, device=DEV), 5, 9))
    assert not torch.equal(a, ops.noise_normal(torch.empty(1003, device=DEV), 5, 10))
    assert not torch.equal(a, ops.noise_normal(torch.empty(1003, device=DEV), 6, 9))
    # distribution: moments, Kolmogorov-Smirnov against the normal CDF, independence of neighbours and streams
    n = 1 << 24
    z = ops.noise_normal(torch.empty(n, device=DEV), 42, 0).double()
    assert abs(float(z.mean())) < 5 / math.sqrt(n)
    assert abs(float(z.var()) - 1) < 5 * math.sqrt(2 / n)
    assert abs(float((z ** 3).mean())) < 5 * math.sqrt(15 / n)
    assert abs(float((z ** 4).mean()) - 3) < 5 * math.sqrt(96 / n)
    assert 5.0 < float(z.abs().max()) < 7.0
    srt = z.sort().values
    cdf = 0.5 * (1 + torch.erf(srt / math.sqrt(2)))
    emp = (torch.arange(1, n + 1, device=DEV, dtype=torch.float64)) / n
    ks = float(torch.maximum((emp - cdf).abs(), (emp - 1 / n - cdf).abs()).max())
    assert ks < 1.95 / math.sqrt(n)  # 0.1 % critical value
    for lag in (1, 2, 3, 4, 5, 8):
        assert abs(float((z[:-lag] * z[lag:]).mean())) < 5 / math.sqrt(n)
    z2 = ops.noise_normal(torch.empty(n, device=DEV), 42, 1).double()
    assert abs(float((z * z2).mean())) < 5 / math.sqrt(n)


def test_roll_strategy_resample(env):
    """svmpc.py:148-150: the new last action is the last step of a sample of the current prior.  With
    the noise buffer given, the kernel's choice is a deterministic function of it (checked against a
    host restatement); over many draws the components follow the mixture weights (H22 clamp included)."""
    from dust_b200 import ops

    L = env["L"]
    torch.manual_seed(5)
    B, N, H, A = 3, 6, 7, 2
    D = H * A
    theta, mu = torch.randn(B, N, H, A), torch.randn(B, N, H, A) * 3
    mix = torch.rand(B, N) + 0.05
    mix[1, 2] = 0.0                       # a zero weight is clamped to float eps, practically never drawn
    var = torch.tensor([4.0, 0.25]).repeat(H)
    noise = torch.randn(B, N, A + 1)
    out = ops.svmpc_forward(cu(torch.zeros(B, N)), cu(theta), cu(mu), cu(mix), cu(1.0 / var), 0.0,
                            roll_strategy=L.ROLL_RESAMPLE, resample_noise=cu(noise))
    nxt = out["theta_next"].cpu()
    assert torch.equal(nxt[:, :, :-1], theta[:, :, 1:])      # shifted along time
    pm = (mix / mix.sum(1, keepdim=True)).clamp(1.19209290e-07, 1 - 1.19209290e-07)
    pm = pm / pm.sum(1, keepdim=True)
    u = 0.5 * torch.erfc(-noise[..., A].double() / math.sqrt(2.0))
    cdf = pm.double().cumsum(1)
    for b in range(B):
        for n in range(N):
            comp = min(int((cdf[b] < u[b, n]).sum()), N - 1)
            exp = mu[b, comp, -1] + var[-A:].sqrt() * noise[b, n, :A]
            assert rel_max(nxt[b, n, -1], exp) <= 1e-6, (b, n, comp)
    with pytest.raises(Exception):
        ops.svmpc_forward(cu(torch.zeros(B, N)), cu(theta), cu(mu), cu(mix), cu(1.0 / var), 0.0,
                          roll_strategy=L.ROLL_RESAMPLE)
    # distribution of the chosen component over many instances, identical centres per component
    Bn = 4096
    mu2 = torch.arange(N, dtype=torch.float32).reshape(1, N, 1, 1).expand(Bn, N, H, A).contiguous() * 100.0
    mix2 = torch.tensor([1.0, 2.0, 3.0, 0.0, 2.0, 2.0]).expand(Bn, N).contiguous()
    nz = ops.noise_normal(torch.empty(Bn, N, A + 1, device=DEV), 3, 1)
    o2 = ops.svmpc_forward(cu(torch.zeros(Bn, N)), cu(torch.zeros(Bn, N, H, A)), cu(mu2), cu(mix2), cu(1.0 / var), 0.0,
                           roll_strategy=L.ROLL_RESAMPLE, resample_noise=nz)
    comp = (o2["theta_next"][:, :, -1, 1] / 100.0).round().long().clamp(0, N - 1).flatten().cpu()   # sigma 0.5 on column 1
    freq = torch.bincount(comp, minlength=N).double() / comp.numel()
    assert float((freq - mix2[0].double() / 10.0).abs().max()) < 0.01
    # the reference-shaped core draws its own noise and stays reproducible
    from dust_b200.inference.core import SvmpcCore
    mk = lambda: SvmpcCore(env["spec"]["particle"], cu(theta), cu(mu), cu(mix), torch.tensor([4.0, 0.25]), torch.ones(A),
                           roll_strategy="resample", seed=11)
    c1, c2 = mk(), mk()
    ll = cu(torch.zeros(B, N))
    c1.forward_step(ll); c2.forward_step(ll)
    assert torch.equal(c1.theta, c2.theta)
    first = c1.theta.clone()
    c1.forward_step(ll)
    assert not torch.equal(c1.theta[:, :, -1], first[:, :, -1])
