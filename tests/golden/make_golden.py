"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The reference has no tests and no golden vectors of its own, so the
fixtures are outputs of the reference itself, run on CPU through oracle/refshim.py.
The reference draws its own random numbers; every tensor it draws
(`torch.distributions` standard-normal draws and `params_dist.sample`) is RECORDED and
stored next to the outputs, so the oracle and the CUDA path are fed the identical noise.

Fixtures that pass through the gpytorch / KDEpy stand-ins are tagged "shim-dependent"
in MANIFEST.json (the stand-ins restate documented third-party behaviour).
"""
import hashlib
import json
import os
import sys
from copy import deepcopy

import numpy as np
import torch
import torch.distributions as dist
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refshim import REFERENCE_ROOT, ref_import  # noqa: E402

disco = ref_import("controllers.disco")
likelihoods = ref_import("inference.likelihoods")
svgd_mod = ref_import("inference.svgd")
svmpc_mod = ref_import("inference.svmpc")
mpf_mod = ref_import("inference.mpf")
base_kernels = ref_import("kernels.base_kernels")
composite_kernels = ref_import("kernels.composite_kernels")
pendulum_mod = ref_import("models.pendulum")
particle_mod = ref_import("models.particle")
from gpytorch.kernels import RBFKernel  # noqa: E402  (the shim's stand-in)

torch.set_num_threads(1)
MANIFEST = {}


# ----------------------------------------------------------------------------------
# noise recording
# ----------------------------------------------------------------------------------
class NoiseRecorder:
    """Wraps torch.distributions' `_standard_normal` so every draw is kept."""

    def __init__(self):
        self.draws = []
        self._mods = [torch.distributions.multivariate_normal, torch.distributions.normal]
        self._orig = [m._standard_normal for m in self._mods]

    def __enter__(self):
        def wrapped(shape, dtype, device, _o=self._orig[0]):
            x = _o(shape, dtype, device)
            self.draws.append(x.detach().clone())
            return x

        for m in self._mods:
            m._standard_normal = wrapped
        return self

    def __exit__(self, *exc):
        for m, o in zip(self._mods, self._orig):
            m._standard_normal = o


class RecordingDist:
    """A params_dist that remembers what it sampled (disco.py:168-174 contract)."""

    def __init__(self, d):
        self.d = d
        self.samples = []

    @property
    def event_shape(self):
        return self.d.event_shape

    @property
    def mean(self):
        return self.d.mean

    def sample(self, shape=torch.Size()):
        x = self.d.sample(shape)
        self.samples.append(x.detach().clone())
        return x

    def log_prob(self, x):
        return self.d.log_prob(x)


def save(name, tags=(), **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    MANIFEST[name] = {
        "tags": list(tags),
        "keys": {k: [str(v.dtype), list(v.shape)] for k, v in out.items()},
        "bytes": os.path.getsize(path),
    }
    print(f"  wrote {name}.npz  {os.path.getsize(path)/1024:.1f} KiB")


# ----------------------------------------------------------------------------------
# configs (the reference's own yaml files, unchanged)
# ----------------------------------------------------------------------------------
with open(os.path.join(REFERENCE_ROOT, "demo", "pendulum_config.yaml")) as f:
    PEND = yaml.load(f, yaml.FullLoader)
with open(os.path.join(REFERENCE_ROOT, "demo", "particle_config.yaml")) as f:
    PART = yaml.load(f, yaml.FullLoader)


def pend_inst_cost(states, controls=None, n_pol=1, debug=None):
    # the demo's cost (pendulum_example.py:21-28); lives in the demo script, not the package
    theta, theta_d = states.chunk(2, dim=1)
    return 50.0 * (theta.cos() - 1) ** 2 + 1.0 * theta_d ** 2


def pend_term_cost(states, n_pol=1, debug=None):
    return pend_inst_cost(states).squeeze()


def make_pendulum(seed, kernel="rbf", params_sampling=True, n_pol=None, S=None, H=None, P=None,
                  log_space=False, optimizer_class=torch.optim.SGD, lr=None, svgd_steps=1, **opt_args):
    ep = PEND["exp_params"]
    torch.manual_seed(seed)
    H = H or ep["horizon"]
    N = n_pol or ep["n_particles"]
    S = S or ep["action_samples"]
    P = P or ep["params_samples"]
    A = ep["ctrl_dim"]
    model = pendulum_mod.PendulumModel(uncertain_params=("length", "mass"))
    prior = svgd_mod.get_gmm(
        torch.randn(N, H, A), torch.ones(N), ep["prior_sigma"] ** 2 * torch.eye(A)
    )
    theta0 = prior.sample([N])
    dyn = dist.Independent(dist.Uniform(torch.tensor([0.6, 0.6]), torch.tensor([1.3, 1.3])), 1)
    ctrl = disco.MultiDISCO(
        observation_space=model.observation_space,
        action_space=model.action_space,
        hz_len=H,
        n_policies=N,
        action_samples=S,
        params_samples=P,
        temperature=1 / ep["alpha"],
        a_cov=ep["ctrl_sigma"] ** 2 * torch.eye(A),
        inst_cost_fn=pend_inst_cost,
        term_cost_fn=pend_term_cost,
        params_sampling=params_sampling,
        params_log_space=log_space,
    )
    if kernel == "rbf":
        k = RBFKernel()
    else:
        k = composite_kernels.iid_mp(
            base_kernel=base_kernels.RBF(bandwidth=-1), ctrl_dim=A, indep_controls=True
        )
    lik = likelihoods.ExponentiatedUtility(ep["alpha"], n_samples=S, controller=ctrl, model=model)
    sv = svmpc_mod.SVMPC(
        init_particles=theta0.clone(),
        prior=prior,
        likelihood=lik,
        kernel=k,
        n_particles=N,
        bw_scale=ep["bandwidth_scaling"],
        n_steps=svgd_steps,
        optimizer_class=optimizer_class,
        lr=ep["learning_rate"] if lr is None else lr,
        weighted_prior=ep["weighted_prior"],
        **opt_args,
    )
    state = torch.as_tensor(ep["init_state"]).clone()
    return dict(model=model, ctrl=ctrl, svmpc=sv, dyn=dyn, state=state, prior=prior, cfg=ep)


def make_particle(seed, kernel="rbf", n_pol=None, S=None, H=None, P=None):
    ep, env = PART["exp_params"], PART["env_params"]
    torch.manual_seed(seed)
    H = H or ep["horizon"]
    N = n_pol or ep["n_particles"]
    S = S or ep["action_samples"]
    P = P or ep["params_samples"]
    A = ep["ctrl_dim"]
    prior = svgd_mod.get_gmm(
        torch.randn(N, H, A), torch.ones(N), ep["prior_sigma"] ** 2 * torch.eye(A)
    )
    theta0 = prior.sample([N])
    dyn_prior = dist.Normal(ep["dyn_prior_arg1"], ep["dyn_prior_arg2"])
    sysk = {"uncertain_params": ["mass"], "mass": dyn_prior.mean}
    model = particle_mod.Particle(**env, **sysk)
    ctrl = disco.MultiDISCO(
        model.observation_space,
        model.action_space,
        H,
        N,
        S,
        temperature=1 / ep["alpha"],
        a_cov=ep["ctrl_sigma"] ** 2 * torch.eye(A),
        params_sampling=ep["sampling"],
        params_samples=P,
        params_log_space=ep["mpf_log_space"],
        inst_cost_fn=model.default_inst_cost,
        term_cost_fn=model.default_term_cost,
    )
    if kernel == "rbf":
        k = RBFKernel()
    else:
        k = composite_kernels.iid_mp(
            base_kernel=base_kernels.RBF(bandwidth=-1), ctrl_dim=A, indep_controls=True
        )
    lik = likelihoods.ExponentiatedUtility(ep["alpha"], controller=ctrl, model=model, n_samples=S)
    sv = svmpc_mod.SVMPC(
        init_particles=theta0.clone(),
        prior=prior,
        likelihood=lik,
        kernel=k,
        n_particles=N,
        bw_scale=ep["bandwidth_scaling"],
        n_steps=1,
        optimizer_class=torch.optim.SGD,
        lr=ep["learning_rate"],
        weighted_prior=ep["weighted_prior"],
    )
    state = torch.as_tensor(env["init_state"], dtype=torch.float).clone()
    return dict(model=model, ctrl=ctrl, svmpc=sv, dyn=dyn_prior, state=state, prior=prior,
                cfg=ep, env=env)


def prior_mu(prior):
    return prior.component_distribution.base_dist.loc.detach().clone()


def prior_mix(prior):
    return prior.mixture_distribution.probs.detach().clone()


# ----------------------------------------------------------------------------------
# G0: obstacle map
# ----------------------------------------------------------------------------------
def gen_map():
    print("G0 obstacle map")
    w = make_particle(0)
    m = w["model"].obst_map.map.astype(np.uint8)
    sha = hashlib.sha1(m.tobytes()).hexdigest()
    save(
        "map_grid4x4",
        packed=np.packbits(m.reshape(-1)),
        shape=np.array(m.shape),
        occupied=np.array(int(m.sum())),
        c_offset=w["model"].obst_map.c_offset.numpy(),
        cell_size=np.array(w["model"].obst_map.cell_size),
    )
    MANIFEST["map_grid4x4"]["sha1_uint8"] = sha
    # a couple of other presets, to pin the generator (obstacle_map.py:101-224)
    for preset, width in [("staggered_3-2-3", 2.0), ("grid_3x3", 1.5), ("single_centred", 3.0),
                          ("staggered_4-3-4-3-4", 1.0), ("grid_6x6", 1.3)]:
        om = ref_import("utils.obstacle_map")
        params = om.get_obst_preset(preset, width)
        mm = om.generate_obstacle_map([22, 22], params, 0.1, map_type="direct")
        arr = mm.map.astype(np.uint8)
        save(f"map_{preset}", packed=np.packbits(arr.reshape(-1)), shape=np.array(arr.shape),
             occupied=np.array(int(arr.sum())), width=np.array(width))
    # collision lookups at awkward coordinates (obstacle_map.py:64-93)
    torch.manual_seed(5)
    X = torch.cat([
        (torch.rand(4000, 2) - 0.5) * 26.0,                     # includes out-of-map points
        torch.tensor([[-11.0, -11.0], [10.99, 10.99], [11.0, 11.0], [0.0, 0.0], [-0.05, 0.05],
                      [4.95, 4.95], [4.949999, 7.05], [7.05, 7.0500001], [-1e9, 1e9]]),
        (torch.randint(-115, 115, (2000, 2)).float() / 10.0),   # exactly on cell boundaries
    ])
    c = w["model"].obst_map.get_collisions(X)
    save("map_collisions", X=X, coll=c)


# ----------------------------------------------------------------------------------
# G1/G2: MultiDISCO.forward (+ step) on recorded noise
# ----------------------------------------------------------------------------------
def gen_forward(kind, seed, sub_s=8, **kw):
    w = make_pendulum(seed, **kw) if kind == "pendulum" else make_particle(seed, **kw)
    ctrl, model = w["ctrl"], w["model"]
    theta = w["svmpc"].theta.detach().clone()
    # a state away from the initial one so clamps / obstacles are exercised
    torch.manual_seed(1000 + seed)
    if kind == "pendulum":
        state = torch.tensor([3.0, 0.0]) + torch.randn(2) * torch.tensor([0.5, 2.0])
    else:
        state = torch.tensor([-9.0, -9.0, 0.0, 0.0]) + torch.rand(4) * torch.tensor([9.0, 9.0, 2.0, 2.0])
    S, N, H, A = ctrl.n_actions, ctrl.n_pol, ctrl.hz_len, ctrl.dim_a
    sigma = ctrl.a_dist.covariance_matrix.diag().sqrt()
    eps = torch.randn(S, N, H, A)
    actions = theta + sigma * eps
    if kind == "pendulum":
        pd = RecordingDist(w["dyn"])
    else:
        # the particle demo samples masses from the MPF prior (log-space GMM over 50 particles)
        x = dist.Normal(2.0, 0.1).sample([50, 1]).clamp(min=1e-6).log()
        mix = dist.Categorical(torch.ones(50))
        comp = dist.Independent(dist.MultivariateNormal(loc=x, covariance_matrix=0.04 * torch.eye(1)), 0)
        pd = RecordingDist(dist.MixtureSameFamily(mix, comp))
    c = deepcopy(ctrl)
    a_mat0 = torch.randn(N, H, A)
    c.a_mat = a_mat0.clone()
    costs, states, acts, weights, plogp = c.forward(state, model, pd, actions)
    out = dict(
        state=state, theta=theta, eps=eps, sigma=sigma, actions=actions,
        params=pd.samples[0] if pd.samples else np.zeros((0,)),
        costs=costs, weights=weights, states_sub=states[:, :sub_s],
        states_last=states[..., -1, :],
        a_mat0=a_mat0, a_mat1=c.a_mat, a_mix1=c.a_mix,
        params_log_p=plogp if plogp is not None else np.zeros((0,)),
        temp=np.array(c.temp), H=np.array(H), N=np.array(N), S=np.array(S), A=np.array(A),
        P=np.array(c.n_params), log_space=np.array(bool(c._params_log_space)),
    )
    # stand-alone controller step (disco.py:396-417)
    for strat in ("argmax", "average"):
        c2 = deepcopy(c)
        nxt = c2.step(strategy=strat, steps=1)
        out[f"step_{strat}_action"] = nxt.clone()
        out[f"step_{strat}_a_seq"] = c2.a_seq
        out[f"step_{strat}_a_mat"] = c2.a_mat
    return out


def gen_forward_all():
    print("G1/G2 MultiDISCO.forward")
    for seed in (0, 1, 2):
        save(f"fwd_pendulum_s{seed}", **gen_forward("pendulum", seed))
        save(f"fwd_particle_s{seed}", **gen_forward("particle", seed))
    # P=1 / no parameter sampling (batched-pendulum shape: N=8,S=256,H=20)
    save("fwd_pendulum_nops", **gen_forward("pendulum", 3, params_sampling=None, n_pol=8, S=256, H=20))
    # internal action sampling (ext_actions=None, disco.py:157-160)
    w = make_pendulum(7)
    c = deepcopy(w["ctrl"])
    c.a_mat = torch.randn(c.n_pol, c.hz_len, c.dim_a)
    a_mat0 = c.a_mat.clone()
    pd = RecordingDist(w["dyn"])
    with NoiseRecorder() as rec:
        costs, states, acts, weights, plogp = c.forward(w["state"], w["model"], pd)
    save("fwd_pendulum_internal", state=w["state"], eps=rec.draws[0], a_mat0=a_mat0,
         params=pd.samples[0], costs=costs, weights=weights, a_mat1=c.a_mat, a_mix1=c.a_mix,
         sigma=c.a_dist.covariance_matrix.diag().sqrt())


# ----------------------------------------------------------------------------------
# G3: SVMPC.optimize + SVMPC.forward, one control step, teacher-forced sequence
# ----------------------------------------------------------------------------------
def run_svmpc_steps(w, n_steps, plant_params=None, with_mpf=None, tag_kernel="rbf"):
    """Runs the reference closed loop and records, per step, inputs and outputs."""
    sv, model = w["svmpc"], w["model"]
    state = w["state"].clone()
    rec_steps = []
    mpf = with_mpf
    dyn = RecordingDist(mpf.prior if mpf is not None else w["dyn"])
    for step in range(n_steps):
        pre = dict(
            state=state.clone(), theta0=sv.theta.detach().clone(), mu0=prior_mu(sv.prior),
            mix0=prior_mix(sv.prior),
        )
        if mpf is not None:
            dyn.d = mpf.prior  # particle_example.py:171 captures once; the object tracks mpf.x by view
            pre["mpf_x0"] = mpf.x.detach().clone()
        dyn.samples.clear()
        with NoiseRecorder() as rec:
            sv.optimize(state, dyn)
        eps = [d for d in rec.draws if d.ndim == 4][0]
        pre["eps"] = eps
        pre["params"] = dyn.samples[0].clone() if dyn.samples else torch.zeros(0)
        post = dict(
            costs=sv.likelihood.last_costs.detach().clone(),
            log_l=sv.likelihood.log_prob(sv.likelihood.last_costs).detach().clone(),
            theta1=sv.theta.detach().clone(),
            phi=-sv.theta.grad.detach().clone(),
        )
        a_seq, p_w = sv.forward(state, dyn)
        post.update(
            a_seq=a_seq.clone(), p_weights=p_w.clone(), i_star=np.array(int(p_w.argmax())),
            theta2=sv.theta.detach().clone(), mu2=prior_mu(sv.prior), mix2=prior_mix(sv.prior),
        )
        action = a_seq[0]
        # plant = the model's own step with "true" parameters (gym is absent)
        if plant_params is None:
            nxt = model.step(state.view(1, -1), action.view(1, -1)).view(-1)
        else:
            nxt = model.step(state.view(1, -1), action.view(1, -1), plant_params).view(-1)
        post["next_state"] = nxt.clone()
        if mpf is not None:
            bw_arg = w.get("mpf_bw")
            gn, bw = mpf.optimize(action.squeeze(), nxt.clone(), bw=bw_arg, n_steps=w["mpf_steps"])
            post["mpf_x1"] = mpf.x.detach().clone()
            post["mpf_grad_norms"] = gn.clone()
            post["mpf_bw"] = np.array(float(bw))
        state = nxt
        rec_steps.append((pre, post))
    return rec_steps


def flatten_steps(rec_steps):
    out = {}
    for i, (pre, post) in enumerate(rec_steps):
        for k, v in pre.items():
            out[f"t{i}_in_{k}"] = v
        for k, v in post.items():
            out[f"t{i}_out_{k}"] = v
    out["n_steps"] = np.array(len(rec_steps))
    return out


def gen_svmpc():
    print("G3 SVMPC step sequences")
    # pendulum, shipped kernel (gpytorch RBFKernel stand-in) -> shim-dependent
    w = make_pendulum(0, kernel="rbf")
    pp = {"length": torch.tensor([[1.1]]), "mass": torch.tensor([[0.8]])}
    save("svmpc_pendulum_rbf", tags=["shim-dependent:gpytorch", "shim-dependent:KDEpy(dead value)"],
         sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(),
         **flatten_steps(run_svmpc_steps(w, 6, plant_params=pp)))
    # pendulum, the reference's own kernels (message passing)
    w = make_pendulum(1, kernel="mp")
    save("svmpc_pendulum_mp", tags=["shim-dependent:KDEpy(dead value)"],
         sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(),
         **flatten_steps(run_svmpc_steps(w, 4, plant_params=pp)))
    w = make_particle(0, kernel="rbf")
    save("svmpc_particle_rbf", tags=["shim-dependent:gpytorch", "shim-dependent:KDEpy(dead value)"],
         sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(),
         **flatten_steps(run_svmpc_steps(w, 4)))
    w = make_particle(1, kernel="mp")
    save("svmpc_particle_mp", tags=["shim-dependent:KDEpy(dead value)"],
         sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(),
         **flatten_steps(run_svmpc_steps(w, 3)))


# ----------------------------------------------------------------------------------
# G4: full dual loop (SVMPC + MPF), as the demos run it
# ----------------------------------------------------------------------------------
def gen_dual():
    print("G4 dual DuSt-MPC loops")
    # pendulum: linear space, explicit bandwidth (avoids the KDEpy stand-in) and Silverman
    for name, bw, tags in (("dual_pendulum_bw", 0.1, ["shim-dependent:gpytorch"]),
                           ("dual_pendulum_silverman", None,
                            ["shim-dependent:gpytorch", "shim-dependent:KDEpy"])):
        w = make_pendulum(2, kernel="rbf")
        ep = w["cfg"]
        torch.manual_seed(42)
        x0 = w["dyn"].sample([ep["mpf_n_particles"]])
        lik = likelihoods.GaussianLikelihood(
            initial_obs=w["state"].clone(), obs_std=ep["mpf_obs_std"],
            model=pendulum_mod.PendulumModel(uncertain_params=("length", "mass")), log_space=False)
        mpf = mpf_mod.MPF(init_particles=x0, likelihood=lik, optimizer_class=torch.optim.SGD,
                          lr=ep["mpf_learning_rate"], bw=bw, bw_scale=ep["mpf_bandwidth_scaling"])
        w["mpf_bw"], w["mpf_steps"] = bw, ep["mpf_steps"]
        prior_bw = float(mpf.prior.component_distribution.base_dist.covariance_matrix[0, 0, 0].sqrt())
        pp = {"length": torch.tensor([[1.1]]), "mass": torch.tensor([[0.8]])}
        save(name, tags=tags, sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(),
             mpf_prior_bw0=np.array(prior_bw), obs_std=np.array(ep["mpf_obs_std"]),
             mpf_lr=np.array(ep["mpf_learning_rate"]),
             **flatten_steps(run_svmpc_steps(w, 4, plant_params=pp, with_mpf=mpf)))
    # particle: log space, bw = 0.5 (particle_config.yaml:31-35)
    w = make_particle(2, kernel="rbf")
    ep = w["cfg"]
    torch.manual_seed(43)
    x0 = w["dyn"].sample([ep["mpf_n_particles"], 1]).clamp(min=1e-6).log()
    lik = likelihoods.GaussianLikelihood(initial_obs=w["state"].clone(), obs_std=ep["mpf_obs_std"],
                                         model=w["model"], log_space=True)
    init_bw = (2 * ep["dyn_prior_arg2"]) ** 1 / 2  # particle_example.py:145 (precedence quirk: = 0.1)
    mpf = mpf_mod.MPF(init_particles=x0, likelihood=lik, optimizer_class=torch.optim.SGD,
                      lr=ep["mpf_learning_rate"], bw=init_bw, bw_scale=ep["mpf_bandwidth_scaling"])
    w["mpf_bw"], w["mpf_steps"] = ep["mpf_bandwidth"], ep["mpf_steps"]
    pp = {"mass": torch.tensor([[3.0]])}
    save("dual_particle", tags=["shim-dependent:gpytorch"],
         sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(),
         mpf_prior_bw0=np.array(init_bw), obs_std=np.array(ep["mpf_obs_std"]),
         mpf_lr=np.array(ep["mpf_learning_rate"]),
         **flatten_steps(run_svmpc_steps(w, 4, plant_params=pp, with_mpf=mpf)))


# ----------------------------------------------------------------------------------
# G5: MPF.optimize stand-alone (one-step likelihood gradient by autograd, mpf.py:40-62)
# ----------------------------------------------------------------------------------
def gen_mpf():
    print("G5 MPF.optimize")
    torch.manual_seed(11)
    # pendulum, linear space, two parameters (length, mass)
    model = pendulum_mod.PendulumModel(uncertain_params=("length", "mass"))
    x0 = dist.Uniform(torch.tensor([0.6, 0.6]), torch.tensor([1.3, 1.3])).sample([50])
    obs0 = torch.tensor([2.5, -1.0])
    lik = likelihoods.GaussianLikelihood(initial_obs=obs0, obs_std=0.1, model=model, log_space=False)
    mpf = mpf_mod.MPF(init_particles=x0.clone(), likelihood=lik, optimizer_class=torch.optim.SGD,
                      lr=1e-3, bw=0.1, bw_scale=1.0)
    action = torch.tensor(1.3)
    true = {"length": torch.tensor([[0.9]]), "mass": torch.tensor([[1.2]])}
    obs1 = model.step(obs0.view(1, -1), action.view(1, -1), true).view(-1)
    # one phi evaluation (for the unit-level check), then the 20-step optimise
    lik.condition(action, obs1)
    phi0 = mpf.phi(0.1)
    lik.loc = obs0  # undo so that optimise() conditions again from the same past
    del lik.past_obs
    lik.condition(action=None, new_obs=obs0)
    gn, bw = mpf.optimize(action, obs1, bw=0.1, n_steps=20)
    save("mpf_pendulum", x0=x0, obs0=obs0, obs1=obs1, action=action, phi0=phi0, x1=mpf.x,
         grad_norms=gn, bw=np.array(bw), prior_bw=np.array(0.1), obs_std=np.array(0.1),
         lr=np.array(1e-3), mu1=mpf.prior.component_distribution.base_dist.loc)
    # particle, log space, one parameter (mass)
    w = make_particle(3)
    model = w["model"]
    x0 = dist.Normal(2.0, 0.1).sample([50, 1]).clamp(min=1e-6).log()
    obs0 = torch.tensor([-8.0, -7.5, 1.0, 2.0])
    lik = likelihoods.GaussianLikelihood(initial_obs=obs0, obs_std=0.1, model=model, log_space=True)
    mpf = mpf_mod.MPF(init_particles=x0.clone(), likelihood=lik, optimizer_class=torch.optim.SGD,
                      lr=0.01, bw=0.1, bw_scale=1.0)
    action = torch.tensor([6.0, -3.0])
    obs1 = model.step(obs0.view(1, -1), action.view(1, -1), {"mass": torch.tensor([[3.0]])}).view(-1)
    gn, bw = mpf.optimize(action, obs1, bw=0.5, n_steps=20)
    save("mpf_particle", x0=x0, obs0=obs0, obs1=obs1, action=action, x1=mpf.x, grad_norms=gn,
         bw=np.array(bw), prior_bw=np.array(0.1), obs_std=np.array(0.1), lr=np.array(0.01))


# ----------------------------------------------------------------------------------
# G6: generic SVGD.phi + bw_median; RBF.eval; iid_mp.eval
# ----------------------------------------------------------------------------------
def gen_phi():
    print("G6 SVGD.phi / kernels")
    s = svgd_mod.SVGD()
    for N in (64, 257, 1024):
        torch.manual_seed(N)
        d = 40
        scale = torch.arange(1, d + 1).float() / 10.0 if N == 257 else torch.ones(d)
        X = (torch.randn(N, d) * scale).requires_grad_(True)
        log_p = lambda x: -0.5 * (x ** 2).sum(-1)  # noqa: E731  standard-normal target
        bw = svgd_mod.bw_median(X, X)
        phi = s.phi(X, log_p, bw)
        d2 = svgd_mod.squared_distance(X.detach(), X.detach())
        med = torch.median(d2)
        save(f"phi_svgd_N{N}", X=X.detach(), score=-X.detach(), bw=bw.detach(), phi=phi.detach(),
             median_d2=med)
    # the reference's own kernels on a tiny set
    torch.manual_seed(3)
    X = torch.randn(6, 80) * 3
    k = base_kernels.RBF(bandwidth=-1)
    K, dK = k.eval(X, X.clone())
    h, d2 = k.compute_bandwidth(X, X.clone())
    mp = composite_kernels.iid_mp(base_kernel=base_kernels.RBF(bandwidth=-1), ctrl_dim=2)
    Kmp, dKmp = mp.eval(X, X.clone())
    save("kernels_small", X=X, K=K, dK=dK, h=np.array(float(h)), d2=d2, Kmp=Kmp, dKmp=dKmp)


# ----------------------------------------------------------------------------------
# G7: pathwise gradient (svmpc.py:58-60 alternative), autograd through the reference rollout
# ----------------------------------------------------------------------------------
def gen_pathwise():
    print("G7 pathwise (autograd) likelihood gradient")
    for kind, seed in (("pendulum", 4), ("particle", 4), ("particle", 5)):
        w = make_pendulum(seed) if kind == "pendulum" else make_particle(seed)
        sv = w["svmpc"]
        torch.manual_seed(2000 + seed)
        if kind == "pendulum":
            state = torch.tensor([2.0, 1.0])
        elif seed == 4:
            state = torch.tensor([-9.0, -9.0, 0.0, 0.0])  # free space (demo start)
        else:
            state = torch.tensor([-7.6, -6.0, 1.0, 0.5])  # drifting towards an obstacle face
        pd = RecordingDist(w["dyn"]) if kind == "pendulum" else RecordingDist(dist.Normal(0.7, 0.1))
        # NB: for the particle the controller is in log space: masses = exp(N(0.7,0.1))
        x = sv.theta.detach().clone().requires_grad_(True)
        with NoiseRecorder() as rec:
            costs, actions = sv.likelihood.sample(x, state, pd)
        log_l = sv.likelihood.log_prob(costs)
        g = torch.autograd.grad(log_l.sum(), x)[0]
        eps = [d for d in rec.draws if d.ndim == 4][0]
        save(f"pathwise_{kind}_s{seed}", state=state, theta=x.detach(), eps=eps,
             params=pd.samples[0], costs=costs.detach(), log_l=log_l.detach(), grad=g,
             sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt())


# ----------------------------------------------------------------------------------
# G8: a whole episode through the reference's own driver (dust/utils/simulations.py:197-260)
# ----------------------------------------------------------------------------------
def gen_episode():
    print("G8 run_particle_episode (reference driver)")
    sims = ref_import("utils.simulations")
    w = make_particle(3, kernel="mp", n_pol=4, S=32, H=20, P=2)
    sv, model, ctrl = w["svmpc"], w["model"], w["ctrl"]
    theta0, mu0 = sv.theta.detach().clone(), prior_mu(sv.prior)
    dyn = RecordingDist(w["dyn"])
    plant_states = []
    orig_cost = ctrl.inst_cost_fn

    def spy_cost(states, *a, **k):   # the driver evaluates the cost of every new plant state (simulations.py:237)
        if states.ndim == 2 and states.shape[0] == 1 and not a and not k:
            plant_states.append(states.detach().clone().view(-1))
        return orig_cost(states, *a, **k)

    ctrl._BaseController__inst_cost_fn = spy_cost   # read-only property over a name-mangled attribute (controllers/base.py:54-62)
    steps, warm_up, load = 8, 2, 1.0
    with NoiseRecorder() as rec:
        cum = sims.run_particle_episode(w["state"].clone(), model, dyn, ctrl, use_svmpc=True, warm_up=warm_up,
                                        svmpc=sv, load=load, steps=steps)
    eps = [d for d in rec.draws if d.ndim == 4]
    assert len(eps) == steps and len(plant_states) == steps, (len(eps), len(plant_states))
    save("episode_particle_mp", tags=["shim-dependent:KDEpy(dead value)"],
         sigma=ctrl.a_dist.covariance_matrix.diag().sqrt(), theta0=theta0, mu0=mu0, init_state=w["state"],
         eps=torch.stack(eps), params=torch.stack([x.reshape(-1) for x in dyn.samples]),
         plant_states=torch.stack(plant_states), cum_cost=np.array(float(cum)), theta_end=sv.theta.detach().clone(),
         steps=np.array(steps), warm_up=np.array(warm_up), load=np.array(load),
         n_param_draws=np.array(len(dyn.samples)))


# ----------------------------------------------------------------------------------
# G9: the next rows of SURVEY 8(f): control regulariser (ctrl_penalty != 1, disco.py:334-344) and
# sigma-point rollouts (MerweScaledUTF + MultiDISCO._sigma_rollout, disco.py:211-292, 312-323)
# ----------------------------------------------------------------------------------
def gen_widen():
    print("G9 control regulariser / sigma-point rollouts")
    utf_mod = ref_import("utils.utf")
    # --- ctrl_penalty != 1: stand-alone controller, two consecutive forward + step calls (a_seq != 0 on the second)
    for kind, seed, pen in (("pendulum", 11, 0.7), ("particle", 12, 0.25)):
        w = make_pendulum(seed) if kind == "pendulum" else make_particle(seed)
        ref, model = w["ctrl"], w["model"]
        kw = dict(observation_space=model.observation_space, action_space=model.action_space, hz_len=ref.hz_len,
                  n_policies=ref.n_pol, action_samples=ref.n_actions, temperature=ref.temp, ctrl_penalty=pen,
                  a_cov=ref.a_dist.covariance_matrix, inst_cost_fn=ref.inst_cost_fn, term_cost_fn=ref.term_cost_fn,
                  params_sampling=True, params_samples=ref.n_params, params_log_space=ref._params_log_space)
        c = disco.MultiDISCO(**kw)
        S, N, H, A = c.n_actions, c.n_pol, c.hz_len, c.dim_a
        torch.manual_seed(2000 + seed)
        c.a_mat = 0.5 * torch.randn(N, H, A)
        if kind == "pendulum":
            pd = RecordingDist(w["dyn"])
            state = torch.tensor([2.5, -0.7])
        else:
            x = dist.Normal(2.0, 0.1).sample([50, 1]).clamp(min=1e-6).log()
            comp = dist.Independent(dist.MultivariateNormal(loc=x, covariance_matrix=0.04 * torch.eye(1)), 0)
            pd = RecordingDist(dist.MixtureSameFamily(dist.Categorical(torch.ones(50)), comp))
            state = torch.tensor([-6.0, -7.0, 0.5, 0.8])
        out = dict(state=state, a_mat0=c.a_mat.clone(), ctrl_penalty=np.array(pen), temp=np.array(c.temp),
                   sigma=c.a_dist.covariance_matrix.diag().sqrt(), log_space=np.array(bool(c._params_log_space)))
        for it in range(2):
            actions = c.a_mat.unsqueeze(0) + out["sigma"] * torch.randn(S, N, H, A)
            out[f"a_seq_in{it}"] = c.a_seq.clone()
            out[f"a_mat_in{it}"] = c.a_mat.clone()
            costs, states, acts, weights, plogp = c.forward(state, model, pd, actions)
            out.update({f"actions{it}": actions, f"params{it}": pd.samples[it], f"costs{it}": costs, f"weights{it}": weights,
                        f"a_mat_fwd{it}": c.a_mat.clone(), f"a_mix{it}": c.a_mix.clone()})
            out[f"action{it}"] = c.step(strategy="average", steps=1).clone()
        save(f"ctrlpen_{kind}", **out)
    # --- sigma points: the demo's transformer (pendulum_config.yaml utf: n=2, alpha=0.5) on the pendulum with a
    # GMM belief (variance.diag() branch) and an MVN belief (covariance_matrix branch); n_pol = 1 as in the demo
    # and n_pol = 3 (the reference's state relabelling, costs unaffected)
    tf = utf_mod.MerweScaledUTF(n=PEND["utf"]["n"], alpha=PEND["utf"]["alpha"])
    sig0 = tf.compute_sigma_points(torch.tensor([0.9, 1.2]), torch.tensor([[0.04, 0.01], [0.01, 0.09]]))
    ut_mean, ut_cov = tf.unscented_transform(sig0)   # NOT the input covariance: the points use rows of the upper factor
    save("utf_points", loc_weights=tf.loc_weights, cov_weights=tf.cov_weights,
         mean=torch.tensor([0.9, 1.2]), cov=torch.tensor([[0.04, 0.01], [0.01, 0.09]]), sigmas=sig0, ut_mean=ut_mean, ut_cov=ut_cov)
    for name, n_pol, belief in (("utf_pendulum_n1_gmm", 1, "gmm"), ("utf_pendulum_n3_mvn", 3, "mvn")):
        w = make_pendulum(21 + n_pol, n_pol=n_pol, S=32, H=12)
        ref, model = w["ctrl"], w["model"]
        c = disco.MultiDISCO(observation_space=model.observation_space, action_space=model.action_space, hz_len=ref.hz_len,
                             n_policies=n_pol, action_samples=ref.n_actions, temperature=ref.temp,
                             a_cov=ref.a_dist.covariance_matrix, inst_cost_fn=ref.inst_cost_fn, term_cost_fn=ref.term_cost_fn,
                             params_sampling=tf, params_log_space=False)
        S, N, H, A = c.n_actions, c.n_pol, c.hz_len, c.dim_a
        torch.manual_seed(3000 + n_pol)
        c.a_mat = 0.5 * torch.randn(N, H, A)
        a_mat0 = c.a_mat.clone()
        if belief == "gmm":   # the demo's dynamics prior: 4 components, sigma 0.1 (pendulum_example.py)
            locs = torch.tensor(PEND["exp_params"]["params_prior_loc"], dtype=torch.float)
            comp = dist.Independent(dist.Normal(locs, PEND["exp_params"]["params_prior_sigma"]), 1)
            pd = dist.MixtureSameFamily(dist.Categorical(torch.ones(4)), comp)
            mean, cov = pd.mean, pd.variance.diag()
        else:
            mean, cov = torch.tensor([0.9, 1.2]), torch.tensor([[0.04, 0.01], [0.01, 0.09]])
            pd = dist.MultivariateNormal(mean, cov)
        state = torch.tensor([2.8, 0.4])
        actions = c.a_mat.unsqueeze(0) + 2.0 * torch.randn(S, N, H, A)
        costs, states, acts, weights, plogp = c.forward(state, model, pd, actions)
        save(name, state=state, a_mat0=a_mat0, actions=actions, mean=mean, cov=cov, belief_is_mvn=np.array(belief == "mvn"),
             locs=locs if belief == "gmm" else np.zeros((0,)), comp_sigma=np.array(PEND["exp_params"]["params_prior_sigma"]),
             loc_weights=tf.loc_weights, costs=costs, weights=weights, states=states, acts_shape=np.array(acts.shape),
             params_log_p=plogp, a_mat1=c.a_mat, a_mix1=c.a_mix, temp=np.array(c.temp),
             sigma=c.a_dist.covariance_matrix.diag().sqrt(), action_avg=deepcopy(c).step(strategy="average").clone())


# ----------------------------------------------------------------------------------
# G10: SVMPC with a non-SGD optimiser (svgd.py:115: Adam is the SVGD default), two SVGD steps per control step
# ----------------------------------------------------------------------------------
def gen_optim():
    print("G10 SVMPC with Adam / momentum SGD")
    pp = {"length": torch.tensor([[1.1]]), "mass": torch.tensor([[0.8]])}
    for name, kw in (("svmpc_pendulum_adam", dict(optimizer_class=torch.optim.Adam, lr=0.05)),
                     ("svmpc_pendulum_momentum", dict(optimizer_class=torch.optim.SGD, lr=0.5, momentum=0.9))):
        w = make_pendulum(31, kernel="rbf", n_pol=4, S=32, H=10, P=3, svgd_steps=2, **kw)
        sv, model = w["svmpc"], w["model"]
        state = w["state"].clone()
        dyn = RecordingDist(w["dyn"])
        out = dict(sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(), theta_init=sv.theta.detach().clone(),
                   mu_init=prior_mu(sv.prior), prior_var=np.array(PEND["exp_params"]["prior_sigma"] ** 2), lr=np.array(kw["lr"]))
        n_ctrl = 3
        for t in range(n_ctrl):
            dyn.samples.clear()
            with NoiseRecorder() as rec:
                sv.optimize(state, dyn)
            eps = [d for d in rec.draws if d.ndim == 4]
            assert len(eps) == 2 and len(dyn.samples) == 2
            out[f"t{t}_state"] = state.clone()
            out[f"t{t}_eps"] = torch.stack(eps)
            out[f"t{t}_params"] = torch.stack([p.clone() for p in dyn.samples])
            out[f"t{t}_theta1"] = sv.theta.detach().clone()
            out[f"t{t}_costs"] = sv.likelihood.last_costs.detach().clone()
            a_seq, p_w = sv.forward(state, dyn)
            out[f"t{t}_a_seq"], out[f"t{t}_p_weights"] = a_seq.clone(), p_w.clone()
            out[f"t{t}_i_star"] = np.array(int(p_w.argmax()))
            out[f"t{t}_theta2"] = sv.theta.detach().clone()
            state = model.step(state.view(1, -1), a_seq[0].view(1, -1), pp).view(-1)
        out["n_ctrl"] = np.array(n_ctrl)
        save(name, tags=["shim-dependent:gpytorch", "shim-dependent:KDEpy(dead value)"], **out)
    # forward(fast_pred=False): the likelihood is sampled again at the updated particles (svmpc.py:135-136)
    w = make_pendulum(33, kernel="rbf", n_pol=4, S=32, H=10, P=3)
    sv, model = w["svmpc"], w["model"]
    state = w["state"].clone()
    dyn = RecordingDist(w["dyn"])
    out = dict(sigma=w["ctrl"].a_dist.covariance_matrix.diag().sqrt(), theta_init=sv.theta.detach().clone(),
               mu_init=prior_mu(sv.prior), prior_var=np.array(PEND["exp_params"]["prior_sigma"] ** 2),
               lr=np.array(PEND["exp_params"]["learning_rate"]))
    for t in range(2):
        dyn.samples.clear()
        with NoiseRecorder() as rec:
            sv.optimize(state, dyn)
            peek = sv.get_weights(state, dyn, fast_pred=False)
            a_seq, p_w = sv.forward(state, dyn, fast_pred=False)
        eps = [d for d in rec.draws if d.ndim == 4]
        assert len(eps) == 3 and len(dyn.samples) == 3      # optimise, get_weights, forward
        out[f"t{t}_state"], out[f"t{t}_eps"] = state.clone(), torch.stack(eps)
        out[f"t{t}_params"] = torch.stack([p.clone() for p in dyn.samples])
        out[f"t{t}_peek"], out[f"t{t}_a_seq"], out[f"t{t}_p_weights"] = peek.clone(), a_seq.clone(), p_w.clone()
        out[f"t{t}_costs_fwd"] = sv.likelihood.last_costs.detach().clone()
        out[f"t{t}_theta2"] = sv.theta.detach().clone()
        state = model.step(state.view(1, -1), a_seq[0].view(1, -1), pp).view(-1)
    out["n_ctrl"] = np.array(2)
    save("svmpc_pendulum_slow_pred", tags=["shim-dependent:gpytorch", "shim-dependent:KDEpy(dead value)"], **out)


# ----------------------------------------------------------------------------------
# G11: MPF with a non-SGD optimiser (mpf.py:23, 59-62: the optimiser is built once and keeps its state across
# optimize() calls; Adam is the SVGD default, svgd.py:115)
# ----------------------------------------------------------------------------------
def gen_mpf_optim():
    print("G11 MPF with Adam / momentum SGD")
    for name, kw in (("mpf_particle_adam", dict(optimizer_class=torch.optim.Adam, lr=0.01)),
                     ("mpf_particle_momentum", dict(optimizer_class=torch.optim.SGD, lr=0.01, momentum=0.9))):
        w = make_particle(5)
        model = w["model"]
        torch.manual_seed(17)
        x0 = dist.Normal(2.0, 0.1).sample([50, 1]).clamp(min=1e-6).log()
        obs0 = torch.tensor([-8.0, -7.5, 1.0, 2.0])
        lik = likelihoods.GaussianLikelihood(initial_obs=obs0, obs_std=0.1, model=model, log_space=True)
        mpf = mpf_mod.MPF(init_particles=x0.clone(), likelihood=lik, bw=0.1, bw_scale=1.0, **kw)
        out = dict(x0=x0, obs0=obs0, prior_bw=np.array(0.1), obs_std=np.array(0.1), lr=np.array(kw["lr"]), bw=np.array(0.5))
        obs = obs0
        for c, action in enumerate((torch.tensor([6.0, -3.0]), torch.tensor([-2.0, 5.0]))):   # two calls: the state carries over
            nxt = model.step(obs.view(1, -1), action.view(1, -1), {"mass": torch.tensor([[3.0]])}).view(-1)
            gn, bw = mpf.optimize(action, nxt, bw=0.5, n_steps=10)
            out[f"c{c}_action"], out[f"c{c}_obs1"], out[f"c{c}_x1"], out[f"c{c}_grad_norms"] = action, nxt, mpf.x.detach().clone(), gn
            obs = nxt
        save(name, **out)


# ----------------------------------------------------------------------------------
# G12: the skid-steer robot's step (skid_steer_robot.py:73-122).  The cart-pole's step raises AttributeError in the
# reference (cartpole.py:150-155), so no fixture can be recorded for it.
# ----------------------------------------------------------------------------------
def gen_aux_models():
    print("G12 skid-steer step")
    sk = ref_import("models.skid_steer_robot")
    torch.manual_seed(23)
    m = sk.SkidSteerRobot(delta_t=0.1)
    M = 64
    states = torch.randn(M, 5) * torch.tensor([3.0, 3.0, 2.0, 0.3, 0.3])
    actions = torch.randn(M, 2) * 0.6            # some beyond the +-0.5 wheel-speed box
    nxt_default = m.step(states, actions, None)
    pd = {"x_icr": 0.1 + 0.2 * torch.rand(M, 1), "wheel_radius": 0.05 + 0.03 * torch.rand(M, 1),
          "axial_distance": 0.4 + 0.2 * torch.rand(M, 1)}
    nxt_sampled = m.step(states, actions, pd)
    try:
        ref_import("models.cartpole").CartPoleModel().step(torch.zeros(1, 4), torch.zeros(1, 1))
        cart_err = ""
    except Exception as e:  # noqa: BLE001
        cart_err = type(e).__name__
    save("skid_steer_step", states=states, actions=actions, next_default=nxt_default, next_sampled=nxt_sampled,
         x_icr=pd["x_icr"], wheel_radius=pd["wheel_radius"], axial_distance=pd["axial_distance"], dt=np.array(0.1),
         cartpole_reference_raises=np.array(int(cart_err == "AttributeError")))


GENERATORS = dict(aux=gen_aux_models, mpf_optim=gen_mpf_optim, map=gen_map, forward=gen_forward_all, svmpc=gen_svmpc, dual=gen_dual, mpf=gen_mpf, phi=gen_phi,
                  pathwise=gen_pathwise, episode=gen_episode, widen=gen_widen, optim=gen_optim)

if __name__ == "__main__":
    # no arguments: everything; otherwise only the named groups, merged into the existing manifest
    picked = sys.argv[1:] or list(GENERATORS)
    mpath = os.path.join(HERE, "MANIFEST.json")
    if sys.argv[1:] and os.path.isfile(mpath):
        MANIFEST.update(json.load(open(mpath))["fixtures"])
    for g in picked:
        GENERATORS[g]()
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "torch": torch.__version__,
                   "fixtures": MANIFEST}, f, indent=1, sort_keys=True)
    total = sum(v["bytes"] for v in MANIFEST.values())
    print(f"total {total/1e6:.2f} MB in {len(MANIFEST)} fixtures")
