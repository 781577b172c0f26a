"""MPF: SVGD particle filter over dynamics parameters (API of dust/inference/mpf.py:12-86); all
`n_steps` updates of one `optimize` call run inside a single kernel launch."""
import numpy as np
import torch
import torch.distributions as dist

from .. import ops
from .belief import ParticleBelief
from .likelihoods import GaussianLikelihood
from .svgd import SVGD


def silvermans_rule(data):
    """Silverman's rule as KDEpy 1.1.0 `bw_selection.silvermans_rule` computes it for 1-D data
    (the reference calls it on the flattened particles, mpf.py:72)."""
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


def bw_silverman(x, bw_scale=1.0):
    """svgd.py:55-81 (+ _select_sigma :10-25): 0.9 * min(std, IQR/1.349) * n^(-1/5); returns a
    per-dimension tensor when the IQR test fails, as the reference does."""
    xn = x.detach().cpu().double().numpy()
    iqr = (np.percentile(xn, 75) - np.percentile(xn, 25)) / 1.349
    std = x.detach().cpu().std(dim=0)
    A = torch.as_tensor(iqr, dtype=torch.float32) if (iqr > 0 and iqr < float(std.min())) else std
    return bw_scale * (0.9 * A * len(x) ** (-0.2))


class MPF(SVGD):
    def __init__(self, init_particles, likelihood: GaussianLikelihood, bw=None, device="cuda", **kwargs):
        super().__init__(**kwargs)
        assert init_particles.ndim == 2, "Particles must be two dimension with batch on dim 0."
        self.device = torch.device(device)
        self.x = torch.as_tensor(init_particles, dtype=torch.float32).to(self.device).contiguous()
        self.likelihood = likelihood
        plain_sgd = self.optimizer_class is torch.optim.SGD and not any(
            self.opt_args.get(k) for k in ("momentum", "weight_decay", "nesterov", "dampening"))
        # plain SGD (the demos) is fused: all n_steps of an optimize() call run inside one launch.  Any other torch
        # optimiser (mpf.py:23: built once, its state lives across calls; Adam is the SVGD default) takes phi from the
        # kernel one step at a time and updates the particles in place on the device (mpf.py:59-62)
        self.optimizer = None if plain_sgd else self.optimizer_class(params=[self.x], **self.opt_args)
        self._spec = None
        self.update_prior(bw)

    def update_prior(self, bw):
        """mpf.py:26-38: GMM with one component per particle, covariance bw^2 I.  Its centres
        alias `self.x` (the kernel updates x in place, exactly like the reference's SGD).  `bw` may be a python
        number, a per-dimension tensor, or the 1-element DEVICE tensor of the Silverman kernel (no host copy)."""
        n, d = self.x.shape
        if bw is None:
            bw = bw_silverman(self.x.flatten(1, -1), self.bw_scale)
        if not torch.is_tensor(bw) and getattr(self, "_prior_bw_key", None) == (float(bw), d):
            self._prior_obj = None       # same scalar bandwidth as last time (the particle configuration): tensors stand
            return
        bw_t = torch.as_tensor(bw, dtype=torch.float32).reshape(-1).to(self.device)
        self._prior_var = (bw_t ** 2).expand(d).clone() if bw_t.numel() == 1 else (bw_t ** 2).clone()   # on the device
        self._prior_inv_var = (1.0 / self._prior_var).contiguous()
        self._prior_bw_key = (float(bw), d) if not torch.is_tensor(bw) else None
        self._prior_obj = None

    @property
    def prior(self):
        """The belief as the controller consumes it: a `ParticleBelief` (samples / log-density with a few device ops;
        any other `torch.distributions` attribute is served by the real MixtureSameFamily, built on demand)."""
        if self._prior_obj is None:
            self._prior_obj = ParticleBelief(self.x, self._prior_var)
        return self._prior_obj

    def optimize(self, action, new_obs, bw=None, n_steps=100, debug=False):
        lik = self.likelihood
        if new_obs is not None:
            lik.condition(action, new_obs)
        assert lik.past_action is not None, \
            "Previous action is None. Need at least one observation to start sampling."
        if self._spec is None:
            self._spec = lik.model.device_spec(device=self.device)
        dev = self.device
        inv_var = self._prior_inv_var
        if bw is None:
            # Silverman's rule on the flattened particles (mpf.py:72), on the device: the kernel reads the bandwidth
            # from device memory, nothing travels to the host
            bw, next_inv_var = ops.silverman_bandwidth(self.x, self.bw_scale, dp=self.x.shape[1])
        else:
            next_inv_var = None
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(1, -1).to(dev).contiguous()  # noqa: E731
        if self.optimizer is None:
            gn = ops.mpf_optimize(self._spec, self.x.unsqueeze(0), f(lik.past_obs), f(lik.past_action), f(lik.loc),
                                  inv_var, lik.sigma, bw, self.opt_args.get("lr", 1e-3), n_steps, lik.log_space)
        else:
            obs0, act, obs1 = f(lik.past_obs), f(lik.past_action), f(lik.loc)
            phi = torch.empty_like(self.x).unsqueeze(0)
            norms = []
            for _ in range(n_steps):
                norms.append(ops.mpf_optimize(self._spec, self.x.unsqueeze(0), obs0, act, obs1, inv_var, lik.sigma, bw, 0.0, 1,
                                              lik.log_space, phi_out=phi)[0])
                self.optimizer.zero_grad()
                self.x.grad = -phi[0]
                self.optimizer.step()
                self.x.grad = None
            gn = torch.cat(norms).unsqueeze(0) if norms else torch.empty(1, 0, device=dev)
        if next_inv_var is not None:
            self._prior_inv_var, self._prior_var, self._prior_obj = next_inv_var, 1.0 / next_inv_var, None
            self._prior_bw_key = None
        else:
            self.update_prior(bw)
        return gn[0], bw
