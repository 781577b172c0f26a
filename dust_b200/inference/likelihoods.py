"""Likelihood objects (API of dust/inference/likelihoods.py:11-135)."""
import math

import torch

from .. import _lib as L
from .. import ops
from .belief import resolve


class GaussianLikelihood:
    """Observation model of the parameter filter: N(new_obs; f(past_obs, past_action; params),
    obs_std^2 I).  `MPF` reads its state; the one-step model evaluation and its parameter
    Jacobian run inside the MPF kernel."""

    def __init__(self, initial_obs, obs_std, model, log_space=False):
        initial_obs = torch.as_tensor(initial_obs, dtype=torch.float32)
        assert initial_obs.ndim == 1, "Gaussian likelihood needs a single dimensional loc tensor."
        self.dim = initial_obs.shape[0]
        self.sigma = obs_std
        self.condition(new_obs=initial_obs, action=None)
        self.model = model
        self.log_space = log_space

    def condition(self, action, new_obs, covariance_matrix=None):
        self.past_obs = getattr(self, "loc", None)
        self.loc = torch.as_tensor(new_obs, dtype=torch.float32)
        self.past_action = action
        if covariance_matrix is not None:
            self.covariance_matrix = covariance_matrix

    def sample(self, theta):
        assert self.past_action is not None, \
            "Previous action is None. Need at least one observation to start sampling."
        params = theta.exp() if self.log_space else theta
        params_dict = self.model.params_to_dict(params)
        states = self.past_obs.reshape(1, -1).repeat(theta.shape[0], 1)
        return self.model.step(states, self.past_action, params_dict)

    def log_prob(self, samples):
        loc = self.loc.to(samples.device).reshape(1, -1)
        d = samples - loc
        lp = -0.5 * (d * d).sum(-1) / self.sigma ** 2 - self.dim * (math.log(self.sigma) + 0.5 * math.log(2 * math.pi))
        return lp.unsqueeze(-1)


class CostLikelihood:
    kind = None

    def __init__(self, n_samples, controller, model):
        self.n_samples = n_samples
        self._last_states = None
        self.last_actions = None
        self.last_policies = None
        self.last_costs = None
        self.params = None
        self._params_log_p = None
        self.controller = controller
        self.model = model
        self._last = {}
        # action noise: the library's counter-based generator (one Philox stream per draw), keyed by
        # torch's seed at construction so that torch.manual_seed still fixes a run.  `noise_fn(shape)`,
        # if set, supplies the standard-normal tensor instead (replay of recorded draws).
        self.noise_fn = None
        self._noise_seed = int(torch.initial_seed()) & (2 ** 63 - 1)
        self._noise_draws = 0

    @property
    def params_log_p(self):
        """Log-density of the last parameter draws under the belief they came from (formed on first access)."""
        self._params_log_p = resolve(self._params_log_p)
        return self._params_log_p

    @params_log_p.setter
    def params_log_p(self, value):
        self._params_log_p = value

    @property
    def last_states(self):
        """Rollout states of the last evaluation [P,S,N,H+1,ds].  SVMPC stores a thunk: the states are rolled out
        again from that step's inputs when (and only when) somebody reads them."""
        if callable(self._last_states):
            self._last_states = self._last_states()
        return self._last_states

    @last_states.setter
    def last_states(self, value):
        self._last_states = value

    def draw_noise(self, shape):
        """Standard-normal action noise [S,N,H,A] on the controller's device (the rsample draw of
        likelihoods.py:90)."""
        dev = self.controller.device
        if self.noise_fn is not None:
            return torch.as_tensor(self.noise_fn(tuple(shape)), dtype=torch.float32).to(dev).contiguous()
        eps = ops.noise_normal(torch.empty(tuple(shape), device=dev), self._noise_seed, (1 << 61) + self._noise_draws)
        self._noise_draws += 1
        return eps

    def sample(self, theta, state, params_dist, eps=None, want=("costs", "log_lik", "lik_weights", "grad_lik")):
        """likelihoods.py:81-101: actions = theta + L eps (rsample), rollouts, costs.  `eps`
        ([S,N,H,A] standard normal) may be supplied; otherwise it is drawn on the device."""
        ctrl = self.controller
        dev = ctrl.device
        theta = torch.as_tensor(theta, dtype=torch.float32).to(dev).contiguous()
        if eps is None:
            eps = self.draw_noise((self.n_samples,) + tuple(theta.shape))
        else:
            eps = torch.as_tensor(eps, dtype=torch.float32).to(dev).contiguous()
        res = ctrl.evaluate(state, self.model, params_dist, eps, theta=theta, want=want, likelihood=self.kind,
                            alpha=self.alpha)
        self._last = res
        self.last_costs = res["costs"]
        self.last_states = res.get("states")
        self.last_actions = theta + ctrl._sigma * eps
        self.params_log_p = res["params_log_p"]
        return self.last_costs, self.last_actions

    def log_prob(self, costs=None):
        if costs is None or costs is self.last_costs:
            if "log_lik" in self._last:
                return self._last["log_lik"]
            costs = self.last_costs
        return self._log_prob(costs)


class ExpectedCost(CostLikelihood):
    kind = L.LIK_EXPECTED_COST

    def __init__(self, alpha, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def _log_prob(self, costs):
        return -self.alpha * costs.mean(dim=0)


class ExponentiatedUtility(CostLikelihood):
    kind = L.LIK_EXP_UTILITY

    def __init__(self, alpha, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def _log_prob(self, costs):
        return (-self.alpha * costs).logsumexp(0) - math.log(costs.size(0))
