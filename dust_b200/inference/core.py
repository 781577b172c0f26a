"""Batched engine behind SVMPC: B independent MPC instances advance together, one kernel
launch per stage.  The drop-in `SVMPC` class is this engine with B = 1.

Stage order per control step (dust/inference/svmpc.py:87-126, 172-200):
  K3 prior score -> K1 rollout/cost/likelihood/analytic gradient [-> K2 adjoint] ->
  K5 phi + SGD update -> K7 weights / argmax / shift / prior refresh.
"""
import math

import torch

from .. import _lib as L
from .. import ops

GPYTORCH_DEFAULT_LENGTHSCALE = math.log(2.0)


class SvmpcCore:
    def __init__(self, spec, theta, mu, mix, prior_var, sigma, alpha=1.0, temperature=1.0, lr=1.0,
                 kernel="gpytorch", lengthscale=GPYTORCH_DEFAULT_LENGTHSCALE, bw_scale=1.0,
                 likelihood=L.LIK_EXP_UTILITY, grad="analytic", roll_strategy="repeat", weighted_prior=False,
                 aliased=False, seed=0, sharded=None, optimizer=None, ops_module=None):
        """theta, mu [B,N,H,A]; mix [B,N]; prior_var [A] (diagonal, shared by all components);
        sigma [A].  sharded: a `distributed.ShardedRollout` -- ONE instance (B = 1) whose parameter draws are
        split over the ranks of its group (every rank holds the same particles and noise).
        optimizer: None = plain SGD, fused into the phi kernel (and into the one-launch step); otherwise a
        callable `params -> torch.optim.Optimizer` (svgd.py:115-124: Adam by default in the reference): phi comes
        from the kernel and the optimiser applies `grad = -phi` to the particles in place (svmpc.py:92-94).
        Its state is keyed by the particle tensor, which `forward_step` replaces -- so, as in the reference
        (svmpc.py:144,158), moment estimates start afresh after every roll.
        ops_module: stand-in for `dust_b200.ops` (host-logic tests run this class on CPU with oracle-backed ops)."""
        self._ops_module = ops_module      # None: the library (kept out of __dict__ as a module so deepcopy works)
        self.spec = spec
        self.theta, self.mu, self.mix = theta.contiguous(), mu.contiguous(), mix.contiguous()
        self.B, self.N, self.H, self.A = theta.shape
        self.D = self.H * self.A
        dev = theta.device
        self.sigma = torch.as_tensor(sigma, dtype=torch.float32).to(dev).contiguous()
        self.set_prior_var(prior_var)
        self.alpha, self.temperature, self.lr = float(alpha), float(temperature), float(lr)
        if kernel not in ("gpytorch", "mp"):
            raise NotImplementedError(f"kernel mode {kernel!r}")
        self.kernel, self.lengthscale, self.bw_scale = kernel, float(lengthscale), float(bw_scale)
        self.likelihood = likelihood
        if grad not in ("analytic", "pathwise"):
            raise ValueError(grad)
        self.grad = grad
        if roll_strategy not in ("repeat", "mean", "resample"):
            raise ValueError("{} is an invalid roll strategy.".format(roll_strategy))
        self.roll_strategy = {"repeat": L.ROLL_REPEAT, "mean": L.ROLL_MEAN, "resample": L.ROLL_RESAMPLE}[roll_strategy]
        # 'resample' draws the new last action from the prior with the library's generator
        self.seed, self.resample_draws = int(seed), 0
        self.weighted_prior = bool(weighted_prior)
        self.aliased = bool(aliased)
        self.sharded = sharded
        self._make_opt, self._opt = optimizer, None
        self.last = {}
        self._pending = None      # forward results the one-launch step already produced (consumed by forward_step)

    @property
    def ops(self):
        return ops if self._ops_module is None else self._ops_module

    def set_prior_var(self, prior_var):
        dev = self.theta.device
        pv = torch.as_tensor(prior_var, dtype=torch.float32).reshape(-1).cpu()
        if pv.numel() == 1:
            pv = pv.expand(self.A)
        self.prior_var = pv.clone()
        full = pv.repeat(self.H)
        self.inv_var = (1.0 / full).to(dev).contiguous()
        self.log_norm = self.ops.gmm_log_norm(full)

    def _flat(self, t):
        return t.reshape(self.B, self.N, self.D)

    def _one_launch(self, state0, eps, params, tiling, lean=False):
        """The whole SVGD step AND the forward results in one launch (`dust_svmpc_step`: the warp kernel for many
        instances, the cluster kernel for few), when the configuration allows it.  -> outputs or None."""
        if not (self.kernel == "gpytorch" and self.grad == "analytic" and self.roll_strategy != L.ROLL_RESAMPLE
                and self._make_opt is None and self.sharded is None and getattr(self, "_fused_ok", True)):
            return None
        ell2 = self.lengthscale ** 2
        want = ("costs", "log_lik", "theta_out") + (() if lean else ("grad_lik", "phi"))
        try:
            return self.ops.svmpc_step(self.spec, state0, eps, self.theta, self.sigma, self.mu, self.mix, self.inv_var,
                                       self.log_norm, 1.0 / (2.0 * ell2), 1.0 / self.N, -1.0 / ell2, self.lr, params=params,
                                       param_tiling=tiling, likelihood=self.likelihood, alpha=self.alpha,
                                       temperature=self.temperature, aliased=self.aliased, do_forward=True,
                                       roll_strategy=self.roll_strategy, weighted_prior=self.weighted_prior, want=want)
        except NotImplementedError:
            self._fused_ok = False
            return None

    def optimize_step(self, state0, eps, params=None, tiling=L.PARAMS_BLOCKED, want_states=False, lean=False):
        """One SVGD step on the policy particles.  state0 [B,ds], eps [B,S,N,H,A] standard normal,
        params [B,P,dp] | None.  Updates theta in place of the old tensor (new storage).  lean: the one-launch
        step skips the diagnostic outputs (likelihood gradient, phi)."""
        self._pending = None
        if not want_states:
            out = self._one_launch(state0, eps, params, tiling, lean)
            if out is not None:
                # the particles after the update (before the roll); what forward_step would compute rides along
                self.theta = out["theta_out"]
                self._pending = out
                self.last = dict(costs=out["costs"], log_lik=out["log_lik"], grad_lik=out.get("grad_lik"), grad_pri=None,
                                 phi=out.get("phi"), states=None, lik_weights=None, theta1=out["theta_out"])
                return self.last
        mu = self.theta if self.aliased else self.mu
        _, grad_pri = self.ops.gmm(self._flat(self.theta), self._flat(mu), self.mix, self.inv_var, self.log_norm,
                              want_log_prob=False)
        want = ["costs", "log_lik"]
        want.append("grad_lik" if self.grad == "analytic" else "lik_weights")
        if want_states:
            want.append("states")
        if self.sharded is not None and params is not None:
            out = self.sharded.evaluate(self.spec, state0, eps, self.theta, self.sigma, params, tiling, self.likelihood,
                                        self.alpha, self.temperature, grad=self.grad)
            grad_lik = out["grad_lik"]
        else:
            out = self.ops.rollout_cost(self.spec, state0, eps, theta=self.theta, sigma=self.sigma, params=params,
                                   param_tiling=tiling, likelihood=self.likelihood, alpha=self.alpha,
                                   temperature=self.temperature, want=tuple(want))
        if self.sharded is not None and params is not None:
            pass
        elif self.grad == "analytic":
            grad_lik = out["grad_lik"]
        else:
            grad_lik = self.ops.rollout_adjoint(self.spec, state0, eps, out["lik_weights"], theta=self.theta,
                                           sigma=self.sigma, params=params, param_tiling=tiling,
                                           likelihood=self.likelihood, alpha=self.alpha)
        score = grad_lik.reshape(self.B, self.N, self.D) + grad_pri
        x = self._flat(self.theta)
        fused_sgd = self._make_opt is None
        if self.kernel == "gpytorch":
            ell2 = self.lengthscale ** 2
            res = self.ops.svgd_phi(x, score, gamma=1.0 / (2.0 * ell2), c1=1.0 / self.N, c2=-1.0 / ell2, lr=self.lr,
                                    want_update=fused_sgd)
        else:
            res = self.ops.svgd_phi(x, score, per_dim=True, bw_scale=self.bw_scale, lr=self.lr, want_update=fused_sgd)
        if fused_sgd:
            self.theta = res["x_out"].reshape(self.B, self.N, self.H, self.A)
        else:
            if self._opt is None or self._opt.param_groups[0]["params"][0] is not self.theta:
                self._opt = self._make_opt([self.theta])      # new particle tensor (first step, or after a roll)
            self._opt.zero_grad()
            self.theta.grad = -res["phi"].reshape(self.B, self.N, self.H, self.A)
            self._opt.step()
            self.theta.grad = None
        self.last = dict(costs=out["costs"], log_lik=out["log_lik"], grad_lik=grad_lik, grad_pri=grad_pri,
                         phi=res["phi"].reshape(self.B, self.N, self.H, self.A), states=out.get("states"),
                         lik_weights=out.get("lik_weights"))
        return self.last

    def control_step(self, state0, eps, params=None, tiling=L.PARAMS_BLOCKED, want_phi=False):
        """optimize_step + forward_step.  One kernel launch when the shape qualifies for the fused
        instance kernel (B >= 74, H*A <= 32, analytic gradient, fixed-lengthscale kernel), else the
        staged sequence.  -> (a_seq [B,H,A], p_weights [B,N], i_star [B])"""
        self.optimize_step(state0, eps, params, tiling, lean=not want_phi)
        return self.forward_step()

    def likelihood_at_particles(self, state0, eps, params=None, tiling=L.PARAMS_BLOCKED):
        """Fresh rollouts at the CURRENT (updated) particles: what `SVMPC.get_weights(fast_pred=False)` does before
        weighing them (svmpc.py:135-137).  -> log_lik [B,N]; the new costs replace the stored ones."""
        out = self.ops.rollout_cost(self.spec, state0, eps, theta=self.theta, sigma=self.sigma, params=params,
                                    param_tiling=tiling, likelihood=self.likelihood, alpha=self.alpha,
                                    temperature=self.temperature, want=("costs", "log_lik"))
        self.last = dict(self.last, costs=out["costs"], log_lik=out["log_lik"])
        self._pending = None
        return out["log_lik"]

    def draw_resample_noise(self):
        """[B,N,A+1] standard normals for roll strategy 'resample' (stream 2^62 + draw index: disjoint
        from the action-noise streams of the batched controller)."""
        buf = torch.empty(self.B, self.N, self.A + 1, device=self.theta.device)
        self.ops.noise_normal(buf, self.seed, (1 << 62) + self.resample_draws)
        self.resample_draws += 1
        return buf

    def forward_step(self, log_lik=None):
        """Weights, best particle, shift, prior refresh.  -> (a_seq [B,H,A], p_weights [B,N], i_star [B])."""
        if log_lik is None and self._pending is not None:
            p, self._pending = self._pending, None        # the one-launch step already weighed, selected and shifted
            self.theta = p["theta_next"]
            self.mu, self.mix, self.aliased = self.theta, p["mix_next"], True
            return p["a_seq"], p["p_weights"], p["i_star"]
        self._pending = None
        log_lik = self.last["log_lik"] if log_lik is None else log_lik
        mu = self.theta if self.aliased else self.mu
        noise = None if self.roll_strategy != L.ROLL_RESAMPLE else self.draw_resample_noise()
        out = self.ops.svmpc_forward(log_lik, self.theta, mu, self.mix, self.inv_var, self.log_norm,
                                roll_strategy=self.roll_strategy, weighted_prior=self.weighted_prior,
                                resample_noise=noise)
        self.theta = out["theta_next"]
        self.mu = self.theta
        self.mix = out["mix_next"]
        self.aliased = True
        return out["a_seq"], out["p_weights"], out["i_star"]
