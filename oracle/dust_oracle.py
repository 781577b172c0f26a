"""CPU oracle for the DuSt-MPC inner loop -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-CPU restatement (eager, dtype-generic: float32 to mirror the reference,
float64 as "truth") of the reference algorithm for the hot path named in BASELINE.json.
Every function cites the reference file:line it follows (paths relative to
/root/reference).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module; `dust_b200/` never does.

Parity pin: the reference ships NO tests and NO golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by tests/golden/make_golden.py (tests/test_oracle_golden.py).  The fixtures
that route through the gpytorch / KDEpy stand-ins of oracle/refshim.py are tagged
"shim-dependent"; for those the third-party behaviour itself is parity-unpinned.
"""
import math

import numpy as np
import torch

# =====================================================================================
# models
# =====================================================================================


def pendulum_step(x, a, length=None, mass=None, g=9.8, dt=0.05, max_torque=2.0, max_speed=8.0):
    """dust/models/pendulum.py:84-100.  x[...,2], a[...,1]; length/mass broadcastable [...,1]
    tensors (sampled) or None (defaults 1.0, python scalars as in the reference)."""
    th, om = x[..., 0:1], x[..., 1:2]
    m = 1.0 if mass is None else mass
    l = 1.0 if length is None else length  # noqa: E741
    u = a.clamp(min=-max_torque, max=max_torque)
    om = om + dt * (-3 * g / (2 * l) * (th + math.pi).sin() + 3.0 / (m * l ** 2) * u)
    om = om.clamp(-max_speed, max_speed)
    th = th + om * dt
    return torch.cat((th, om), dim=-1)


def pendulum_cost(x):
    """demo/pendulum_example.py:21-28 (instantaneous and terminal are the same function)."""
    th, om = x[..., 0], x[..., 1]
    return 50.0 * (th.cos() - 1) ** 2 + 1.0 * om ** 2


class ParticleCfg:
    """Particle environment constants (dust/models/particle.py:11-106, 292-326)."""

    def __init__(self, grid, dt=0.015, max_speed=5.0, max_accel=10.0, cell_size=0.1,
                 c_offset=(110.0, 110.0), target=(9.0, 9.0, 0.0, 0.0), w_qpos=0.5, w_qvel=0.25,
                 w_ctrl=0.2, w_obs=1.0e6, w_qpos_T=1.0e3, w_qvel_T=0.1, can_crash=True,
                 with_obstacle=True, default_mass=2.0):
        self.grid = torch.as_tensor(np.asarray(grid), dtype=torch.float32)  # [nx, ny] 0/1
        self.dt, self.max_speed, self.max_accel = dt, max_speed, max_accel
        self.cell_size, self.c_offset = cell_size, torch.tensor(c_offset)
        self.target = torch.tensor(target)
        self.w_state = torch.tensor([w_qpos, w_qpos, w_qvel, w_qvel])
        self.w_term = torch.tensor([w_qpos_T, w_qpos_T, w_qvel_T, w_qvel_T])
        self.w_ctrl = torch.tensor([w_ctrl, w_ctrl])
        self.w_obs = w_obs
        self.can_crash, self.with_obstacle = can_crash, with_obstacle
        self.default_mass = default_mass


def collisions(cfg, xy):
    """dust/utils/obstacle_map.py:64-93: floor(x/cell + offset) -> clamp per axis -> gather."""
    occ = (xy * (1 / cfg.cell_size) + cfg.c_offset.to(xy.dtype)).floor()
    ix = occ[..., 0].clamp(0, cfg.grid.shape[0] - 1).long()
    iy = occ[..., 1].clamp(0, cfg.grid.shape[1] - 1).long()
    return cfg.grid.to(xy.dtype)[ix, iy]


def particle_step(cfg, x, a, mass=None):
    """dust/models/particle.py:136-166 (deterministic, control_type='acceleration')."""
    m = cfg.default_mass if mass is None else mass
    u = torch.clamp(a / m, min=-cfg.max_accel, max=cfg.max_accel)
    x_dot = torch.cat((x[..., 2:], u), dim=-1)
    if cfg.can_crash and cfg.with_obstacle:
        c = collisions(cfg, x[..., 0:2]).unsqueeze(-1)
        nx = x + x_dot * cfg.dt * (1 - c)
    else:
        nx = x + x_dot * cfg.dt
    nx = torch.cat((nx[..., :2], nx[..., 2:].clamp(-cfg.max_speed, cfg.max_speed)), dim=-1)
    return nx


def particle_inst_cost(cfg, x, a):
    """dust/models/particle.py:170-198 (raw actions in the control cost)."""
    obst = cfg.w_obs * collisions(cfg, x[..., 0:2]) if cfg.with_obstacle else 0.0
    d = x - cfg.target.to(x.dtype)
    return ((d * d) * cfg.w_state.to(x.dtype)).sum(-1) + ((a * a) * cfg.w_ctrl.to(x.dtype)).sum(-1) + obst


def particle_term_cost(cfg, x):
    """dust/models/particle.py:202-225."""
    obst = cfg.w_obs * collisions(cfg, x[..., 0:2]) if cfg.with_obstacle else 0.0
    d = x - cfg.target.to(x.dtype)
    return ((d * d) * cfg.w_term.to(x.dtype)).sum(-1) + obst


class Model:
    """Bundle of (step, inst cost, term cost, dims) for one of the two shipped systems."""

    def __init__(self, kind, cfg=None):
        self.kind, self.cfg = kind, cfg
        if kind == "pendulum":
            self.ds, self.da, self.dp = 2, 1, 2  # params columns: (length, mass)  pendulum_example.py:165
        elif kind == "particle":
            self.ds, self.da, self.dp = 4, 2, 1  # params column: mass
        else:
            raise ValueError(kind)

    def step(self, x, a, params=None):
        if self.kind == "pendulum":
            if params is None:
                return pendulum_step(x, a)
            return pendulum_step(x, a, length=params[..., 0:1], mass=params[..., 1:2])
        return particle_step(self.cfg, x, a, None if params is None else params[..., 0:1])

    def inst_cost(self, x, a):
        return pendulum_cost(x) if self.kind == "pendulum" else particle_inst_cost(self.cfg, x, a)

    def term_cost(self, x):
        return pendulum_cost(x) if self.kind == "pendulum" else particle_term_cost(self.cfg, x)


# =====================================================================================
# controller: rollout, cost, soft-min weights   (dust/controllers/disco.py)
# =====================================================================================


def tile_params(params, S, N, log_space=False):
    """disco.py:171-179: per-rollout parameters [P,S,N,dp] (see `rollout` for the 1-D quirk)."""
    if params is None:
        return None
    p = params.exp() if log_space else params
    dp = 1 if p.ndim == 1 else p.shape[-1]
    return p.repeat(1, S * N).reshape(params.shape[0], S, N, dp)


def rollout(model, state, actions, params=None, log_space=False):
    """disco.py:139-209.  actions [S,N,H,A]; params [P,dp] as SAMPLED (exponentiated here if
    log_space, disco.py:173-174) or None.  Returns states [P,S,N,H+1,ds].
    Quirk: a scalar-event params_dist samples a 1-D [P] tensor and `repeat(1, S*N)` then tiles
    it INTERLEAVED (rollout row r uses params[r % P]) instead of blocked (params[r // (S*N)])."""
    S, N, H, A = actions.shape
    P = 1 if params is None else params.shape[0]
    R = P * S * N
    p = tile_params(params, S, N, log_space)
    if p is not None:
        p = p.reshape(R, -1)
    acts = actions.reshape(-1, H, A).repeat(P, 1, 1)
    x = state.reshape(1, -1).to(actions.dtype).expand(R, -1)
    states = [x]
    for t in range(H):
        x = model.step(x, acts[:, t], p)
        states.append(x)
    return torch.stack(states, dim=1).reshape(P, S, N, H + 1, model.ds)


def trajectory_costs(model, states, actions, a_reg=0.0, a_seq=None, a_mat=None, a_pre=None):
    """disco.py:294-346: sum_t inst(x_t, a_t), t=0..H-1, + term(x_H); mean over P; + control
    regulariser (zero when ctrl_penalty == 1, disco.py:90)."""
    P, S, N, Hp1, ds = states.shape
    H = Hp1 - 1
    acts = actions.unsqueeze(0).expand(P, -1, -1, -1, -1)
    inst = model.inst_cost(states[..., :-1, :], acts).sum(-1)
    term = model.term_cost(states[..., -1, :])
    cost = (inst + term).mean(0)
    if a_reg != 0.0:
        eps = actions - a_seq
        ctrl = a_reg * torch.einsum("snha,nha->sn", -eps, a_mat @ a_pre)
        cost = cost + ctrl
    return cost


def softmin_update(costs, eps, temp):
    """disco.py:380-393: global-min shift, per-policy log-sum-exp over samples, weights,
    policy-mean increment and mixture weights."""
    beta = costs.min()
    log_costs = -1 * (costs - beta) / temp
    eta = log_costs.logsumexp(0)
    omega = log_costs - eta
    w = omega.exp()
    delta = torch.einsum("sn,snha->nha", w, eps)
    a_mix = (eta - eta.logsumexp(0)).exp()
    return w, delta, a_mix


def disco_forward(model, state, actions, params=None, log_space=False, temp=1.0, a_seq=None,
                  a_reg=0.0, a_mat=None, a_pre=None, eps=None):
    """disco.py:348-394.  With `ext_actions` the perturbation used for the policy update is
    actions - a_seq (disco.py:164); with internal sampling (`eps` given: the a_dist draw,
    actions = eps + a_mat, disco.py:157-160) it is the draw itself.
    Returns dict(costs, states, weights, delta, a_mix)."""
    states = rollout(model, state, actions, params, log_space)
    costs = trajectory_costs(model, states, actions, a_reg, a_seq, a_mat, a_pre)
    if eps is None:
        eps = actions if a_seq is None else actions - a_seq
    w, delta, a_mix = softmin_update(costs, eps, temp)
    return dict(costs=costs, states=states, weights=w, delta=delta, a_mix=a_mix)


def merwe_weights(n, alpha=1e-3, beta=2.0, kappa=0.0):
    """utf.py:81-91 -> (loc_weights, cov_weights), float32 [2n+1]."""
    lam = alpha ** 2 * (n + kappa) - n
    c = 0.5 / (n + lam)
    loc = torch.ones(2 * n + 1, dtype=torch.float) * c
    cov = torch.ones(2 * n + 1, dtype=torch.float) * c
    cov[0] = lam / (n + lam) + (1 - alpha ** 2 + beta)
    loc[0] = lam / (n + lam)
    return loc, cov


def merwe_sigma_points(mu, K, alpha=1e-3, kappa=0.0):
    """utf.py:93-123 -> sigmas [n, 2n+1] (upper Cholesky factor U of (lambda + n) K; mu[i] is added to ROW i)."""
    mu, K = torch.as_tensor(mu, dtype=torch.float), torch.as_tensor(K, dtype=torch.float)
    n = mu.shape[0]
    lam = alpha ** 2 * (n + kappa) - n
    U = torch.linalg.cholesky(((lam + n) * K).transpose(-2, -1)).transpose(-2, -1)
    sig = torch.zeros(n, 2 * n + 1, dtype=torch.float)
    sig[:, 0] = mu
    sig[:, 1:n + 1] = U + mu.view(-1, 1)
    sig[:, n + 1:] = -U + mu.view(-1, 1)
    return sig


def disco_forward_sigma(model, state, actions, sigmas, loc_w, temp=1.0, a_seq=None):
    """MultiDISCO.forward with a MerweScaledUTF transformer (disco.py:211-292, 312-323, 380-393).
    actions [S,N,H,A]; sigmas [n, pts].  Rollout row r = (s*N + n)*pts + k carries action row (s, n) and
    sigma point k.  Cost quirk reproduced: the flat instantaneous costs (index r*H + t) are viewed as
    (-1, pts), i.e. weighted in runs of pts CONSECUTIVE (k, t) entries, not across the sigma points of
    one step; the terminal costs (index r) are grouped by sigma point as intended.
    Returns dict(costs [S,N], states [S*pts, N, H+1, ds], weights, delta, a_mix)."""
    S, N, H, A = actions.shape
    pts = sigmas.shape[1]
    acts = actions.repeat(1, 1, pts, 1).reshape(-1, H, A)          # disco.py:257-259
    prm = sigmas.T.repeat(S * N, 1)                                 # disco.py:262-264
    x = state.reshape(1, -1).to(actions.dtype).expand(S * N * pts, -1)
    states = [x]
    for t in range(H):
        x = model.step(x, acts[:, t], prm)
        states.append(x)
    states = torch.stack(states, dim=1).reshape(S * pts, N, H + 1, model.ds)   # disco.py:281-284 (relabels rows)
    x_vec = states[..., :-1, :].reshape(-1, model.ds)
    x_fin = states[..., -1, :].reshape(-1, model.ds)
    a_vec = actions.reshape(-1, A)     # only its shape matters for the pendulum cost; see below for the particle
    if model.kind == "particle":       # inst_cost_fn(x_vec, a_vec) needs matching rows: the reference would fail here
        raise NotImplementedError("sigma-point costs with an action-dependent cost: shapes differ in the reference")
    inst = model.inst_cost(x_vec, a_vec).reshape(-1, pts) @ loc_w
    term = model.term_cost(x_fin).reshape(-1, pts) @ loc_w
    costs = inst.view(S, N, H).sum(dim=-1) + term.view(S, N)
    eps = actions if a_seq is None else actions - a_seq
    w, delta, a_mix = softmin_update(costs, eps, temp)
    return dict(costs=costs, states=states, weights=w, delta=delta, a_mix=a_mix)


def disco_step(a_mat, a_mix, low, high, strategy="argmax", steps=1):
    """disco.py:396-417.  Returns (next_actions[steps,A], a_seq', a_mat').
    Quirk: with "argmax" the reference's a_seq is a VIEW of a_mat[i*] (integer-like index), so
    the in-place clamp (disco.py:409-410) also clamps that row of a_mat before it is rolled."""
    a_mat = a_mat.clone()
    if strategy == "argmax":
        i = int(a_mix.argmax())
        a_mat[i] = torch.max(torch.min(a_mat[i], high), low)
        a_seq = a_mat[i].clone()
    elif strategy == "average":
        a_seq = torch.einsum("nha,n->ha", a_mat, a_mix)
        a_seq = torch.max(torch.min(a_seq, high), low)
    else:
        raise ValueError("Invalid value for strategy.")
    nxt = a_seq[:steps].clone()
    a_seq = a_seq.roll(-steps, 0)
    a_seq[-steps:] = 0
    a_mat = a_mat.roll(-steps, 1)
    a_mat[:, -steps:] = 0
    return nxt, a_seq, a_mat


# =====================================================================================
# likelihoods and gradients
# =====================================================================================


def exp_utility_log_prob(costs, alpha):
    """likelihoods.py:127-135."""
    return (-alpha * costs).logsumexp(0) - math.log(costs.shape[0])


def expected_cost_log_prob(costs, alpha):
    """likelihoods.py:113-119."""
    return -alpha * costs.mean(0)


def analytic_lik_grad(costs, actions, theta, sigma, alpha):
    """svmpc.py:46-54: sum_s softmax_s(-alpha C[s,i]) (a[s,i] - theta_i) / sigma^2."""
    w = torch.softmax(-costs * alpha, dim=0)
    d_log_pi = (actions - theta) / sigma ** 2
    return (w[..., None, None] * d_log_pi).sum(0)


def pathwise_lik_grad_autograd(model, state, theta, eps, sigma, params, log_space, alpha):
    """svmpc.py:58-60 (the commented alternative): d/dtheta sum_i log_l_i by autograd through
    this oracle's own rollout.  Returns (grad[N,H,A], costs, log_l)."""
    x = theta.detach().clone().requires_grad_(True)
    actions = x + sigma * eps
    out = disco_forward(model, state, actions, params, log_space)
    log_l = exp_utility_log_prob(out["costs"], alpha)
    (g,) = torch.autograd.grad(log_l.sum(), x)
    return g, out["costs"].detach(), log_l.detach()


def pathwise_lik_grad_adjoint(model, state, theta, eps, sigma, params, log_space, alpha,
                              want_param_grad=False, weights=None):
    """Hand-derived reverse-time adjoint of rollout+cost (SURVEY.md §9 "Adjoint equations"),
    the CPU statement of what the CUDA adjoint kernels compute.  Forward: disco.py:139-209,
    294-346; clamp sub-gradients are inclusive (torch.clamp backward, H18); floor/collision
    terms carry no gradient."""
    S, N, H, A = eps.shape
    actions = theta + sigma * eps
    states = rollout(model, state, actions, params, log_space)  # [P,S,N,H+1,ds]
    costs = trajectory_costs(model, states, actions)
    # d log_l_n / d C[s,n] = -alpha w[s,n]; `weights` holds them fixed (the adjoint is linear in them: tests feed
    # the device's own soft-min weights so that cost rounding inside exp(-alpha C) does not dominate)
    w = torch.softmax(-alpha * costs, dim=0) if weights is None else weights.to(costs.dtype)
    P = states.shape[0]
    dt = states.dtype
    a = actions.unsqueeze(0).expand(P, -1, -1, -1, -1)
    gA = torch.zeros(P, S, N, H, A, dtype=dt)
    pv = tile_params(params, S, N, log_space)
    pv = None if pv is None else pv.to(dt)
    gP = None
    if model.kind == "pendulum":
        g_, d_ = 9.8, 0.05
        l = torch.ones(P, 1, 1, dtype=dt) if pv is None else pv[..., 0]  # noqa: E741
        m = torch.ones(P, 1, 1, dtype=dt) if pv is None else pv[..., 1]
        c1, c2 = -3 * g_ / (2 * l), 3.0 / (m * l ** 2)
        thH, omH = states[..., H, 0], states[..., H, 1]
        lam_th = -100.0 * (thH.cos() - 1) * thH.sin()
        lam_om = 2.0 * omH
        gl, gm = torch.zeros_like(lam_th), torch.zeros_like(lam_th)
        for t in range(H - 1, -1, -1):
            th, om, at = states[..., t, 0], states[..., t, 1], a[..., t, 0]
            u = at.clamp(-2.0, 2.0)
            sn = (th + math.pi).sin()
            pre = om + d_ * (c1 * sn + c2 * u)
            m8 = ((pre >= -8.0) & (pre <= 8.0)).to(dt)
            m2 = ((at >= -2.0) & (at <= 2.0)).to(dt)
            gom = lam_om + d_ * lam_th
            g = m8 * gom
            gA[..., t, 0] = g * d_ * c2 * m2
            gl = gl + g * d_ * (3 * g_ / (2 * l ** 2) * sn - 6.0 * u / (m * l ** 3))
            gm = gm + g * d_ * (-3.0 * u / (m ** 2 * l ** 2))
            lam_th = lam_th + g * d_ * c1 * (th + math.pi).cos() - 100.0 * (th.cos() - 1) * th.sin()
            lam_om = g + 2.0 * om
        gP = torch.stack((gl, gm), dim=-1)
    else:
        cfg = model.cfg
        m = torch.full((P, 1, 1), float(cfg.default_mass), dtype=dt) if pv is None else pv[..., 0]
        tgt = cfg.target.to(dt)
        ws, wt, wc = cfg.w_state.to(dt), cfg.w_term.to(dt), cfg.w_ctrl.to(dt)
        xH = states[..., H, :]
        lam = 2.0 * wt * (xH - tgt)  # [P,S,N,4]
        gm = torch.zeros(P, S, N, dtype=dt)
        for t in range(H - 1, -1, -1):
            xt, at = states[..., t, :], a[..., t, :]
            am = at / m.unsqueeze(-1)
            ma = ((am >= -cfg.max_accel) & (am <= cfg.max_accel)).to(dt)
            u = am.clamp(-cfg.max_accel, cfg.max_accel)
            c = collisions(cfg, xt[..., 0:2]) if (cfg.can_crash and cfg.with_obstacle) else 0.0
            k = (cfg.dt * (1 - c)).unsqueeze(-1) if torch.is_tensor(c) else cfg.dt
            vpre = xt[..., 2:] + u * k
            mv = ((vpre >= -cfg.max_speed) & (vpre <= cfg.max_speed)).to(dt)
            gv = mv * lam[..., 2:]
            gA[..., t, :] = k * gv * ma / m.unsqueeze(-1) + 2.0 * wc * at
            gm = gm - (k * gv * ma * at / (m.unsqueeze(-1) ** 2)).sum(-1)
            lam_p = lam[..., :2] + 2.0 * ws[:2] * (xt[..., :2] - tgt[:2])
            lam_v = gv + k * lam[..., :2] + 2.0 * ws[2:] * (xt[..., 2:] - tgt[2:])
            lam = torch.cat((lam_p, lam_v), dim=-1)
        gP = gm.unsqueeze(-1)
    # chain: log_l_n = lse_s(-alpha mean_p C);  a = theta + sigma eps  =>  da/dtheta = I
    coef = (-alpha * w / P)[None, :, :, None, None]
    grad_theta = (coef * gA).sum((0, 1))
    log_l = exp_utility_log_prob(costs, alpha)
    if want_param_grad:
        # d sum_n log_l_n / d (sampled parameter p); blocked tiling only
        assert params is not None and params.ndim == 2
        gp = (-alpha * w)[None, :, :, None] * gP / P
        gp = gp.sum((1, 2))
        if log_space:
            gp = gp * params.exp().to(dt)
        return grad_theta, costs, log_l, gp
    return grad_theta, costs, log_l


# =====================================================================================
# GMM prior   (dust/inference/svgd.py:84-89, mpf.py:26-38)
# =====================================================================================


_F32_EPS = 1.1920928955078125e-07


def gmm_log_mix(mix):
    """torch.distributions.Categorical(probs=mix).logits as MixtureSameFamily uses them:
    probs are normalised, CLAMPED to [eps, 1-eps] (float32 eps, probs_to_logits) and the
    log is re-normalised with log_softmax -- a zero weight becomes log(eps) ~ -15.9, not -inf."""
    probs = mix / mix.sum()
    return torch.log_softmax(probs.clamp(_F32_EPS, 1.0 - _F32_EPS).log(), dim=-1)


def gmm_log_prob(x, mu, mix, var):
    """MixtureSameFamily(Categorical(mix), Independent(MVN(mu, diag(var)), .)).log_prob(x).
    x [M,D], mu [K,D], mix [K] (unnormalised), var [D] or scalar (diagonal covariance)."""
    D = x.shape[-1]
    var = torch.as_tensor(var, dtype=x.dtype).expand(D)
    diff = x[:, None, :] - mu[None, :, :]
    maha = (diff * diff / var).sum(-1)
    comp = -0.5 * (D * math.log(2 * math.pi) + maha) - 0.5 * var.log().sum()
    logpi = gmm_log_mix(mix)
    return (comp + logpi).logsumexp(-1)


def gmm_score(x, mu, mix, var):
    """grad_x sum_i log GMM(x_i) = sum_k r_ik (mu_k - x_i) / var  (svmpc.py:41, mpf.py:45)."""
    D = x.shape[-1]
    var = torch.as_tensor(var, dtype=x.dtype).expand(D)
    diff = x[:, None, :] - mu[None, :, :]
    maha = (diff * diff / var).sum(-1)
    logpi = gmm_log_mix(mix)
    r = torch.softmax(-0.5 * maha + logpi, dim=-1)
    return -(r[..., None] * diff).sum(1) / var


# =====================================================================================
# distances, bandwidths, kernels
# =====================================================================================


def sq_dists_addmm(x, y):
    """svgd.py:28-39 (clamped at zero)."""
    xn = x.pow(2).sum(-1, keepdim=True)
    yn = y.pow(2).sum(-1, keepdim=True)
    return (yn.t() - 2.0 * (x @ y.t()) + xn).clamp(min=0)


def sq_dists_rbf(x, y):
    """kernels/base_kernels.py:58-62 (three products, NOT clamped)."""
    return -2 * (x @ y.t()) + (x * x).sum(-1).unsqueeze(1) + (y * y).sum(-1).unsqueeze(0)


def lower_median(v):
    """torch.median semantics (H4): element of rank (n-1)//2 of the sorted flat array."""
    flat = v.reshape(-1)
    return flat.sort().values[(flat.numel() - 1) // 2]


def bw_median(x, bw_scale=1.0, tol=1e-5):
    """svgd.py:42-52: bw = scale * max(sqrt(med/2) / log(N+1), tol)."""
    med = lower_median(sq_dists_addmm(x, x))
    h = torch.sqrt(0.5 * med) / math.log(x.shape[0] + 1.0)
    return bw_scale * h.clamp(min=tol), med


def rbf_bandwidth(x, y, ell=-1.0, ell_scale=1.0, minimum_bw=1e-5):
    """kernels/base_kernels.py:53-89: h = clamp(scale * (med | ell^2) / log(N+1), min)."""
    d2 = sq_dists_rbf(x, y)
    h = lower_median(d2) if ell < 0 else torch.as_tensor(ell ** 2, dtype=x.dtype)
    h = h / np.log(x.shape[0] + 1)
    h = (ell_scale * h).clamp(min=minimum_bw)
    return h.to(x.dtype), d2


def rbf_eval(x, y, ell=-1.0, ell_scale=1.0):
    """kernels/base_kernels.py:91-108: K = exp(-d2/h), dK[i,j,:] = K_ij (x_i - y_j) 2/h."""
    h, d2 = rbf_bandwidth(x, y, ell, ell_scale)
    K = (-d2 / h).exp()
    dK = K.unsqueeze(2) * (x.unsqueeze(1) - y) * 2 / h
    return K, dK, h


def iid_mp_eval(x, y):
    """kernels/composite_kernels.py:33-64 with indep_controls=True: an independent scalar RBF
    (own median bandwidth) per flattened (timestep, control) column."""
    m, D = x.shape
    K = torch.zeros(m, m, D, dtype=x.dtype)
    dK = torch.zeros(m, m, D, dtype=x.dtype)
    hs = torch.zeros(D, dtype=x.dtype)
    for q in range(D):
        k, dk, h = rbf_eval(x[:, q:q + 1], y[:, q:q + 1])
        K[:, :, q], dK[:, :, q], hs[q] = k, dk.squeeze(2), h
    return K, dK, hs


def phi_unified(x, score, gamma, c1, c2):
    """phi = c1 K S + c2 (rowsum(K) o X - K X), K = exp(-gamma d2)."""
    d2 = sq_dists_addmm(x, x)
    K = (-gamma * d2).exp()
    return c1 * (K @ score) + c2 * (K.sum(1, keepdim=True) * x - K @ x)


def phi_svgd(x, score, bw):
    """svgd.py:127-135 with default_kernel (svgd.py:92-99): canonical SVGD direction."""
    N = x.shape[0]
    return phi_unified(x, score, 1.0 / (2.0 * bw ** 2), 1.0 / N, 1.0 / (N * bw ** 2))


GPYTORCH_DEFAULT_LENGTHSCALE = math.log(2.0)  # softplus(0); svmpc.py:78 never sets it (H1)


def phi_svmpc_gpytorch(x, score, ell=GPYTORCH_DEFAULT_LENGTHSCALE):
    """svmpc.py:76-83 (shipped `kernel: rbf`): attractive, un-normalised kernel gradient (H3)."""
    N = x.shape[0]
    return phi_unified(x, score, 1.0 / (2.0 * ell ** 2), 1.0 / N, -1.0 / ell ** 2)


def phi_mpf(x, score, bw):
    """mpf.py:53-56."""
    N = x.shape[0]
    return phi_unified(x, score, 1.0 / (2.0 * bw ** 2), 1.0 / N, -1.0 / bw ** 2)


def phi_svmpc_iid_mp(x, score):
    """svmpc.py:64-74 with the message-passing kernel."""
    K, dK, _ = iid_mp_eval(x, x.clone())
    return (K * score.unsqueeze(0)).mean(1) + dK.mean(1)


def phi_svmpc_rbf_composed(x, score):
    """svmpc.py:64-74 as it would act with the plain RBF kernel (the in-class call raises, H2)."""
    K, dK, _ = rbf_eval(x, x.clone())
    return (K.unsqueeze(2) * score.unsqueeze(0)).mean(1) + dK.mean(1)


# =====================================================================================
# SVMPC control step   (dust/inference/svmpc.py)
# =====================================================================================


class SvmpcState:
    """theta [N,H,A]; GMM prior (mu, mix, var[A]); `aliased`: after the first update_prior the
    reference's prior centres share storage with theta (get_gmm's x.detach(), svgd.py:88 +
    MultivariateNormal's loc.expand view), so later in-place SGD steps move the centres too."""

    def __init__(self, theta, mu, mix, var, aliased=False):
        self.theta, self.mu, self.mix, self.var, self.aliased = theta, mu, mix, var, aliased


def svmpc_optimize(model, st, state, eps, sigma, params, log_space, alpha, lr, kernel="rbf",
                   lik="exp_utility", grad="analytic"):
    """svmpc.py:87-126 with n_steps=1 and an SGD optimiser.  Returns dict of intermediates;
    updates st.theta (and st.mu when aliased)."""
    N, H, A = st.theta.shape
    x = st.theta.clone()
    flat = lambda t: t.reshape(N, -1)  # noqa: E731
    var_full = torch.as_tensor(st.var, dtype=x.dtype).expand(A).repeat(H)
    mu = st.theta if st.aliased else st.mu
    grad_pri = gmm_score(flat(x), flat(mu), st.mix, var_full).reshape(N, H, A)
    actions = x + sigma * eps  # likelihoods.py:85-90 (rsample with diagonal scale)
    out = disco_forward(model, state, actions, params, log_space)
    costs = out["costs"]
    if lik == "exp_utility":
        log_l = exp_utility_log_prob(costs, alpha)
    else:
        log_l = expected_cost_log_prob(costs, alpha)
    if grad == "analytic":
        grad_lik = analytic_lik_grad(costs, actions, x, sigma, alpha)
    else:
        grad_lik = pathwise_lik_grad_adjoint(model, state, x, eps, sigma, params, log_space, alpha)[0]
    score = grad_lik + grad_pri
    if kernel == "rbf":
        phi = phi_svmpc_gpytorch(flat(x), flat(score)).reshape(N, H, A)
    elif kernel == "mp":
        phi = phi_svmpc_iid_mp(flat(x), flat(score)).reshape(N, H, A)
    else:
        raise ValueError(kernel)
    theta1 = st.theta + lr * phi  # SGD: theta -= lr * (-phi)
    st.theta = theta1
    if st.aliased:
        st.mu = theta1
    return dict(costs=costs, log_l=log_l, grad_lik=grad_lik, grad_pri=grad_pri, phi=phi,
                theta1=theta1, states=out["states"], actions=actions)


def svmpc_forward(st, costs, alpha, weighted_prior, roll_strategy="repeat", lik="exp_utility"):
    """svmpc.py:128-200 with fast_pred=True: weights from the PRE-update costs and the prior
    evaluated at the POST-update particles (H10), argmax, shift, prior refresh."""
    N, H, A = st.theta.shape
    flat = lambda t: t.reshape(N, -1)  # noqa: E731
    var_full = torch.as_tensor(st.var, dtype=st.theta.dtype).expand(A).repeat(H)
    log_l = exp_utility_log_prob(costs, alpha) if lik == "exp_utility" else expected_cost_log_prob(costs, alpha)
    mu = st.theta if st.aliased else st.mu
    log_p = gmm_log_prob(flat(st.theta), flat(mu), st.mix, var_full)
    log_w = log_l + log_p
    p_w = (log_w - log_w.logsumexp(0)).exp()
    i_star = int(p_w.argmax())
    a_seq = st.theta[i_star].clone()
    th = st.theta.roll(-1, dims=-2)
    if roll_strategy == "repeat":
        th[..., -1, :] = th[..., -2, :]
    elif roll_strategy == "mean":
        th[..., -1, :] = th.mean(dim=-2)
    else:
        raise ValueError("{} is an invalid roll strategy.".format(roll_strategy))
    st.theta = th
    st.mu = th
    st.mix = p_w.clone() if weighted_prior else torch.ones_like(p_w)
    st.aliased = True
    return a_seq, p_w, i_star


# =====================================================================================
# MPF: SVGD over dynamics-parameter particles   (dust/inference/mpf.py, likelihoods.py:11-64)
# =====================================================================================


def model_step_param_grad(model, obs0, action, x, log_space):
    """One model step from obs0 for every parameter particle and d s'/d x (closed form of the
    autograd call at mpf.py:50).  Returns (s' [Np,ds], J [Np,ds,dp])."""
    Np = x.shape[0]
    dt_ = x.dtype
    p = x.exp() if log_space else x
    s0 = obs0.reshape(1, -1).to(dt_).expand(Np, -1)
    a = action.reshape(1, -1).to(dt_).expand(Np, -1)
    s1 = model.step(s0, a, p)
    J = torch.zeros(Np, model.ds, model.dp, dtype=dt_)
    if model.kind == "pendulum":
        g_, d_ = 9.8, 0.05
        l, m = p[:, 0], p[:, 1]  # noqa: E741
        th, om, at = s0[:, 0], s0[:, 1], a[:, 0]
        u = at.clamp(-2.0, 2.0)
        sn = (th + math.pi).sin()
        pre = om + d_ * (-3 * g_ / (2 * l) * sn + 3.0 / (m * l ** 2) * u)
        m8 = ((pre >= -8.0) & (pre <= 8.0)).to(dt_)
        dom_dl = m8 * d_ * (3 * g_ / (2 * l ** 2) * sn - 6.0 * u / (m * l ** 3))
        dom_dm = m8 * d_ * (-3.0 * u / (m ** 2 * l ** 2))
        J[:, 1, 0], J[:, 1, 1] = dom_dl, dom_dm
        J[:, 0, 0], J[:, 0, 1] = d_ * dom_dl, d_ * dom_dm
    else:
        cfg = model.cfg
        m = p[:, 0:1]
        am = a / m
        ma = ((am >= -cfg.max_accel) & (am <= cfg.max_accel)).to(dt_)
        u = am.clamp(-cfg.max_accel, cfg.max_accel)
        c = collisions(cfg, s0[:, 0:2]).unsqueeze(-1) if (cfg.can_crash and cfg.with_obstacle) else 0.0
        k = cfg.dt * (1 - c)
        vpre = s0[:, 2:] + u * k
        mv = ((vpre >= -cfg.max_speed) & (vpre <= cfg.max_speed)).to(dt_)
        J[:, 2:, 0] = mv * k * ma * (-a / m ** 2)
    if log_space:
        J = J * p.unsqueeze(1)
    return s1, J


def mpf_phi(model, x, obs0, action, obs1, obs_std, prior_var, bw, log_space):
    """mpf.py:40-57.  The prior's centres alias self.x (MultivariateNormal's expand view of
    loc, mpf.py:32-38), i.e. they are the CURRENT particles; prior_var [dp] stays as set by the
    last update_prior."""
    Np = x.shape[0]
    ones = torch.ones(Np, dtype=x.dtype)
    grad_prior = gmm_score(x, x, ones, prior_var)
    s1, J = model_step_param_grad(model, obs0, action, x, log_space)
    r = (obs1.reshape(1, -1).to(x.dtype) - s1) / obs_std ** 2
    grad_lik = torch.einsum("is,isp->ip", r, J)
    score = grad_lik + grad_prior
    return phi_mpf(x, score, bw)


def mpf_optimize(model, x, obs0, action, obs1, obs_std, prior_var, bw, lr, n_steps, log_space):
    """mpf.py:64-86 with an SGD optimiser.  Returns (x', grad_norms[n_steps])."""
    x = x.clone()
    norms = []
    for _ in range(n_steps):
        phi = mpf_phi(model, x, obs0, action, obs1, obs_std, prior_var, bw, log_space)
        x = x + lr * phi
        norms.append(phi.norm())
    return x, torch.stack(norms)


def silverman_kdepy(data):
    """KDEpy 1.1.0 bw_selection.silvermans_rule on the flattened particles (mpf.py:72).  The
    third-party source is absent from the container: restated from its documented behaviour."""
    data = np.asarray(data, dtype=np.float64).reshape(-1)
    n = data.shape[0]
    if n == 1:
        return 1.0
    iqr = (np.percentile(data, 75) - np.percentile(data, 25)) / 1.349
    std = np.std(data, ddof=1)
    sigma = min(std, iqr) if iqr > 0 else std
    if sigma > 0:
        return float(sigma * (n * 3 / 4.0) ** (-1 / 5))
    iqr = (np.percentile(data, 99) - np.percentile(data, 1)) / 4.6526957480816815
    return float(iqr * (n * 3 / 4.0) ** (-1 / 5)) if iqr > 0 else 1.0


# =====================================================================================
# large-N phi: tiled float64 truth (the reference cannot run N=65536: K alone is 17 GB)
# =====================================================================================


def phi_unified_tiled(x, score, gamma, c1, c2, tile=2048, dtype=torch.float64):
    x, score = x.to(dtype), score.to(dtype)
    N = x.shape[0]
    out = torch.zeros_like(x)
    xn = (x * x).sum(-1)
    for i0 in range(0, N, tile):
        xi = x[i0:i0 + tile]
        d2 = (xn[i0:i0 + tile, None] + xn[None, :] - 2.0 * xi @ x.t()).clamp(min=0)
        K = (-gamma * d2).exp()
        out[i0:i0 + tile] = c1 * (K @ score) + c2 * (K.sum(1, keepdim=True) * xi - K @ x)
    return out


def median_sq_dist_tiled(x, tile=2048, dtype=torch.float32):
    """Exact lower median of all N^2 clamped squared distances without materialising them:
    two-pass radix select on the float bit pattern (non-negative floats order like ints)."""
    x = x.to(dtype)
    assert dtype == torch.float32
    N = x.shape[0]
    k = (N * N - 1) // 2
    xn = (x * x).sum(-1)

    def tiles():
        for i0 in range(0, N, tile):
            xi = x[i0:i0 + tile]
            d2 = (xn[None, :] - 2.0 * (xi @ x.t()) + xn[i0:i0 + tile, None]).clamp(min=0)
            yield d2.reshape(-1).view(torch.int32).to(torch.int64)

    hist = torch.zeros(1 << 16, dtype=torch.int64)
    for b in tiles():
        hist += torch.bincount(b >> 16, minlength=1 << 16)
    cum = hist.cumsum(0)
    hi = int((cum > k).nonzero()[0])
    below = int(cum[hi - 1]) if hi > 0 else 0
    hist2 = torch.zeros(1 << 16, dtype=torch.int64)
    for b in tiles():
        sel = b[(b >> 16) == hi]
        hist2 += torch.bincount(sel & 0xFFFF, minlength=1 << 16)
    cum2 = hist2.cumsum(0)
    lo = int((cum2 > (k - below)).nonzero()[0])
    bits = (hi << 16) | lo
    return torch.tensor([bits], dtype=torch.int32).view(torch.float32)[0]


# =====================================================================================
# the reference's other two forward models: step only (no cost function / demo ships for them)
# =====================================================================================


def skid_steer_step(x, a, dt, x_icr=0.2, wheel_radius=0.0625, axial_distance=0.475, lo=(-0.5, -0.5), hi=(0.5, 0.5)):
    """dust/models/skid_steer_robot.py:73-122.  x [M,5], a [M,2] -> [M,5]; parameters scalars or [M,1] tensors."""
    px, py, th = x[:, 0:1], x[:, 1:2], x[:, 2:3]
    r = a[:, 0:1].clamp(lo[0], hi[0])
    l = a[:, 1:2].clamp(lo[1], hi[1])  # noqa: E741
    lin = (r + l) * math.pi * wheel_radius
    ang = (r - l) * 2 * math.pi * wheel_radius / axial_distance
    fwd = lin * dt
    lat = -ang * x_icr * dt
    nx = px + fwd * torch.cos(th) - lat * torch.sin(th)
    ny = py + fwd * torch.sin(th) + lat * torch.cos(th)
    return torch.cat([nx, ny, th + ang * dt, lin.expand_as(px), ang.expand_as(px)], dim=1)


def cartpole_step(x, a, dt=0.05, g=9.8, m_c=1.0, m_p=0.1, length=1.0, mu_c=0.5e-3, mu_p=2e-6, f_mag=10.0):
    """dust/models/cartpole.py:147-172, the method body as written (`mass = m_c + m_c`, :160).  The reference's method
    itself raises AttributeError before reaching it (name-mangled `self.__params_dict`, :150-155): PARITY UNPINNED for
    this function -- there is no reference output to check it against, only the source text."""
    px, x_d, th, th_d = x.chunk(4, dim=1)
    acts = torch.clamp(a, min=-1, max=1) * f_mag
    mass = m_c + m_c
    pm = m_p * length
    cart_friction = mu_c * x_d.sign()
    pole_friction = (mu_p * th_d) / pm
    factor = (acts + pm * th.sin() * th_d ** 2 - cart_friction) / mass
    tdd_num = g * th.sin() - th.cos() * factor - pole_friction
    tdd_den = length * (4.0 / 3 - (m_p * th.cos() ** 2) / mass)
    theta_dd = tdd_num / tdd_den
    x_dd = factor - pm * theta_dd * torch.cos(th) / mass
    return x + torch.cat([x_d, x_dd, th_d, theta_dd], dim=1) * dt
