"""The belief over dynamics parameters that `MPF.prior` hands to the controller (dust/inference/mpf.py:26-38): a Gaussian
mixture with one component per parameter particle, equal weights, covariance bw^2 I, whose centres ALIAS the particle
tensor (quirk H24: an object captured once keeps tracking the particles with the covariance it was built with).

The reference builds a `torch.distributions.MixtureSameFamily` for it after every filter update; constructing and
sampling that object costs a few dozen small launches and a host synchronisation per control step.  This class
answers what the control loop asks of it -- `sample`, `log_prob`, `mean`, `variance` -- with a handful of device ops
and no synchronisation, and turns into the real torch object (built once, on demand) for anything else."""
import math

import torch
import torch.distributions as dist


class ParticleBelief:
    def __init__(self, centres, var):
        """centres [n, d] device tensor (kept by reference); var [d] tensor (host or device) of component variances."""
        self._x = centres
        self._var = torch.as_tensor(var, dtype=torch.float32).reshape(-1)
        if self._var.numel() == 1:
            self._var = self._var.expand(centres.shape[1])
        self._std_dev = None
        self._torch = None

    # --- what the control loop uses ----------------------------------------------------------
    @property
    def event_shape(self):
        return torch.Size([self._x.shape[1]])

    @property
    def batch_shape(self):
        return torch.Size([])

    def _std(self):
        if self._std_dev is None:
            self._std_dev = self._var.to(self._x.device).sqrt()
        return self._std_dev

    def sample(self, sample_shape=torch.Size()):
        """Component index uniformly at random, then the component's Gaussian (disco.py:168: `params_dist.sample`)."""
        shape = tuple(sample_shape)
        n, d = self._x.shape
        idx = torch.randint(n, shape, device=self._x.device)
        return self._x[idx] + self._std() * torch.randn(shape + (d,), device=self._x.device)

    rsample = sample

    def log_prob(self, value):
        x, var = self._x, self._var.to(self._x.device)
        v = torch.as_tensor(value, dtype=torch.float32).to(x.device)
        d2 = ((v.unsqueeze(-2) - x) ** 2 / var).sum(-1)
        log_norm = -0.5 * (x.shape[1] * math.log(2 * math.pi) + var.log().sum())
        return torch.logsumexp(-0.5 * d2, dim=-1) - math.log(x.shape[0]) + log_norm

    @property
    def mean(self):
        return self._x.mean(0)

    @property
    def variance(self):
        return self._var.to(self._x.device) + self._x.var(0, unbiased=False)

    # --- everything else: the torch.distributions object the reference would hold ---------------
    def as_torch(self):
        if self._torch is None:
            n = self._x.shape[0]
            cov = torch.diag(self._var.to(self._x.device))
            comp = dist.Independent(dist.MultivariateNormal(loc=self._x, covariance_matrix=cov), reinterpreted_batch_ndims=0)
            self._torch = dist.MixtureSameFamily(dist.Categorical(torch.ones(n, device=self._x.device)), comp)
        return self._torch

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.as_torch(), name)


class Lazy:
    """A value computed on first use (`params_log_p`: the log-density of the parameter draws is stored by the
    reference on every control step but only read by diagnostics)."""

    def __init__(self, fn):
        self._fn, self._val, self._done = fn, None, False

    def get(self):
        if not self._done:
            self._val, self._done, self._fn = self._fn(), True, None
        return self._val


def resolve(v):
    return v.get() if isinstance(v, Lazy) else v
