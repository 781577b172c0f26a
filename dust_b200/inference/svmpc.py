"""SVMPC: Stein-variational MPC over control-sequence particles (API of
dust/inference/svmpc.py:14-200), every stage on the GPU through `SvmpcCore` (B = 1)."""
import torch

from .. import _lib as L
from .. import ops
from ..kernels.base_kernels import RBF
from ..kernels.composite_kernels import iid_mp
from .core import SvmpcCore
from .likelihoods import CostLikelihood
from .svgd import SVGD, get_gmm


def _kernel_mode(kernel):
    if isinstance(kernel, iid_mp):
        if not kernel.indep_controls:
            raise NotImplementedError("iid_mp(indep_controls=False) has no device kernel")
        base = kernel.base_kernel
        # the per-dimension device kernel implements the median heuristic with the default clamp only
        # (base_kernels.py:64-89 with ell < 0, minimum_bw = 1e-5): anything else must not pass silently
        if getattr(base, "ell", -1) >= 0:
            raise NotImplementedError("iid_mp(base_kernel=RBF(bandwidth >= 0)): a fixed bandwidth has no device kernel "
                                      "(use bandwidth=-1, the median heuristic)")
        if float(getattr(base, "minimum_bw", 1e-5)) != 1e-5:
            raise NotImplementedError("iid_mp(base_kernel=RBF(minimum_bw != 1e-5)) has no device kernel")
        return "mp", 0.0, base.ell_scale
    if isinstance(kernel, RBF):
        raise NotImplementedError(
            "a plain RBF kernel inside SVMPC.phi raises in the reference as well (broadcast of k_XX [N,N] "
            "against the score, svmpc.py:68-71); use iid_mp(base_kernel=RBF()) or RBFKernel()")
    if hasattr(kernel, "lengthscale"):  # gpytorch RBFKernel or the bundled stand-in
        return "gpytorch", float(torch.as_tensor(kernel.lengthscale).detach().reshape(-1)[0]), 1.0
    raise NotImplementedError(f"kernel {kernel!r} has no device kernel")


class SVMPC(SVGD):
    def __init__(self, init_particles, prior, likelihood: CostLikelihood, roll_strategy="repeat",
                 weighted_prior=False, grad="analytic", **kwargs):
        super().__init__(**kwargs)
        self.likelihood = likelihood
        ctrl = likelihood.controller
        dev = ctrl.device
        self.device = dev
        self.w_prior = weighted_prior
        self.roll_strategy = roll_strategy
        theta = torch.as_tensor(init_particles, dtype=torch.float32).to(dev)
        self.n_particles = theta.shape[0] if self.n_particles is None else self.n_particles
        comp = prior.component_distribution.base_dist
        mu = comp.loc.detach().to(dev, torch.float32)
        cov = comp.covariance_matrix.detach().to("cpu", torch.float32)
        cov = cov.reshape(-1, cov.shape[-2], cov.shape[-1])[0]
        if float((cov - torch.diag(cov.diag())).abs().max()) != 0.0:
            raise NotImplementedError("only diagonal prior covariances have a device kernel")
        self._prior_cov = cov
        mix = prior.mixture_distribution.probs.detach().to(dev, torch.float32)
        mode, ell, scale = _kernel_mode(self.kernel)
        if getattr(ctrl, "a_reg", 0) != 0 or getattr(ctrl, "_tf", None) is not None:
            raise NotImplementedError("SVMPC drives the rollout kernel directly: a controller with ctrl_penalty != 1 or "
                                      "sigma-point parameter tiling is only supported stand-alone (MultiDISCO.forward)")
        plain_sgd = self.optimizer_class is torch.optim.SGD and not any(
            k in self.opt_args and self.opt_args[k] for k in ("momentum", "weight_decay", "nesterov", "dampening"))
        # plain SGD (the demos) is fused into the update kernels; any other torch optimiser (svgd.py:115: Adam by
        # default) takes phi from the kernel and steps the particles in place on the device (svmpc.py:92-94)
        import functools
        opt = None if plain_sgd else functools.partial(self.optimizer_class, **self.opt_args)
        self._core = SvmpcCore(
            spec=ctrl._spec(likelihood.model), theta=theta.unsqueeze(0), mu=mu.unsqueeze(0), mix=mix.unsqueeze(0),
            prior_var=cov.diag(), sigma=ctrl._sigma, alpha=likelihood.alpha, temperature=ctrl.temp,
            lr=self.opt_args.get("lr", 1e-3), kernel=mode, lengthscale=ell if mode == "gpytorch" else 1.0,
            bw_scale=scale, likelihood=likelihood.kind, grad=grad, roll_strategy=roll_strategy,
            weighted_prior=weighted_prior, optimizer=opt)
        self._prior_obj = prior
        self._prior_stale = False
        self.last_phi = None

    # --- state exposed with the reference's names ------------------------------------------
    @property
    def theta(self):
        return self._core.theta[0]

    @theta.setter
    def theta(self, value):
        self._core.theta = torch.as_tensor(value, dtype=torch.float32).to(self.device).unsqueeze(0).contiguous()
        self._core._pending = None       # forward results computed for the old particles no longer apply

    @property
    def prior(self):
        """The GMM prior as a torch.distributions object (rebuilt lazily after update_prior)."""
        if self._prior_stale:
            c = self._core
            centres = c.theta[0] if c.aliased else c.mu[0]
            self._prior_obj = get_gmm(centres, c.mix[0], self._prior_cov.to(self.device))
            self._prior_stale = False
        return self._prior_obj

    # --- the control step --------------------------------------------------------------------
    def _evaluate(self, state, params_dist, eps=None):
        ctrl, lik, c = self.likelihood.controller, self.likelihood, self._core
        if eps is None:
            eps = lik.draw_noise((lik.n_samples, c.N, c.H, c.A))
        else:
            eps = torch.as_tensor(eps, dtype=torch.float32).to(self.device).contiguous()
        params, tiling, params_log_p = ctrl._sample_params(lik.model, params_dist)
        state0 = torch.as_tensor(state, dtype=torch.float32).reshape(1, -1).to(self.device).contiguous()
        return state0, eps.unsqueeze(0), params, tiling, params_log_p

    def step(self, state, params_dist, bw=None, sigma=None, eps=None):
        """svmpc.py:87-95: theta <- theta + lr * phi (SGD on -phi).  `bw` and `sigma` are accepted for
        signature compatibility: bw is dead in the reference for both shipped kernels (H1) and sigma
        is the controller's."""
        lik, c = self.likelihood, self._core
        state0, eps, params, tiling, params_log_p = self._evaluate(state, params_dist, eps)
        theta_before = c.theta
        out = c.optimize_step(state0, eps, params, tiling)
        lik._last = {"log_lik": out["log_lik"][0]}
        lik.last_costs = out["costs"][0]
        if lik.controller.return_states:
            # the rollouts are only read for rendering (particle_example.py:190): produced on first access, from the
            # inputs of this step, instead of being written on every control step
            spec, sigma = c.spec, c.sigma
            lik.last_states = lambda: ops.rollout_cost(spec, state0, eps, theta=theta_before, sigma=sigma, params=params,
                                                       param_tiling=tiling, want=("costs", "states"))["states"][0]
        else:
            lik.last_states = None
        lik.last_actions = None
        lik._last_eps, lik._last_theta = eps[0], None
        lik.params_log_p = params_log_p
        self.last_phi = out["phi"][0]
        self._prior_stale = self._prior_stale or c.aliased

    def optimize(self, state, params_dist, bw=None, n_steps=None, debug=False, eps=None):
        n_steps = self.n_steps if n_steps is None else n_steps
        for _ in range(n_steps):
            self.step(state, params_dist, bw, None, eps=eps)

    def _resample_likelihood(self, state, params_dist, eps=None):
        """svmpc.py:135-136: new action noise and parameter draws, rollouts at the updated particles."""
        lik, c = self.likelihood, self._core
        state0, eps, params, tiling, params_log_p = self._evaluate(state, params_dist, eps)
        log_lik = c.likelihood_at_particles(state0, eps, params, tiling)
        lik._last = {"log_lik": log_lik[0]}
        lik.last_costs = c.last["costs"][0]
        lik.last_states, lik.last_actions = None, None
        lik.params_log_p = params_log_p
        return log_lik

    def get_weights(self, state, params_dist, fast_pred=True, eps=None):
        """svmpc.py:128-140.  fast_pred re-uses the costs of the optimise step; otherwise the likelihood is
        sampled again at the updated particles (`eps`: optional recorded noise, as in `optimize`)."""
        log_lik = None if fast_pred else self._resample_likelihood(state, params_dist, eps)
        return self._peek_weights(log_lik)

    def _peek_weights(self, log_lik=None):
        import copy
        c = copy.copy(self._core)
        return c.forward_step(log_lik)[1][0]

    def forward(self, state, params_dist, steps=-1, fast_pred=True, eps=None):
        """svmpc.py:172-200 -> (a_seq [H,A], p_weights [N])."""
        if steps != -1:
            raise NotImplementedError("SVMPC.forward: only steps=-1 (shift by one) is supported")
        log_lik = None if fast_pred else self._resample_likelihood(state, params_dist, eps)
        a_seq, p_w, i_star = self._core.forward_step(log_lik)
        self._prior_stale = True
        self.i_star = i_star[0]
        return a_seq[0], p_w[0]

    def update_prior(self, weights=None):
        c = self._core
        w = torch.ones(c.N, device=self.device) if weights is None else torch.as_tensor(weights).to(self.device)
        c.mix = (w if self.w_prior else torch.ones_like(w)).reshape(1, -1).float().contiguous()
        c.mu, c.aliased = c.theta, True
        self._prior_stale = True
