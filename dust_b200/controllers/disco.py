"""MultiDISCO: multi-policy information-theoretic MPC (API of dust/controllers/disco.py:16-417),
with rollout, cost and soft-min weights computed by the K1 CUDA kernels."""
import torch

from ..inference.belief import Lazy, resolve
from .. import _lib as L
from .. import ops
from ..utils.utf import MerweScaledUTF
from .base import BaseController

Empty = torch.Size([])


def diag_sigma(cov):
    """sqrt(diag(cov)) for a diagonal action covariance; anything else has no device kernel."""
    cov = torch.as_tensor(cov, dtype=torch.float32)
    if cov.ndim != 2 or float((cov - torch.diag(cov.diag())).abs().max()) != 0.0:
        raise NotImplementedError("only diagonal action covariances have a device kernel")
    return cov.diag().sqrt()


class MultiDISCO(BaseController):
    def __init__(self, observation_space, action_space, hz_len, n_policies, action_samples, temperature=1.0,
                 ctrl_penalty=1.0, a_cov=None, inst_cost_fn=None, term_cost_fn=None, params_sampling=True,
                 params_samples=4, params_log_space=False, init_actions=None, return_states=True, **kwargs):
        super().__init__(observation_space, action_space, hz_len, inst_cost_fn, term_cost_fn, **kwargs)
        self.n_pol = n_policies
        self.n_actions = action_samples
        self.temp = temperature
        self.a_reg = temperature * (1 - ctrl_penalty)
        dev = self.device
        a_cov = torch.eye(self.dim_a) if a_cov is None else torch.as_tensor(a_cov, dtype=torch.float32)
        self._sigma = diag_sigma(a_cov.cpu()).to(dev)
        self.a_dist = torch.distributions.multivariate_normal.MultivariateNormal(
            torch.zeros(self.dim_a, device=dev), a_cov.to(dev))
        self.a_pre = torch.inverse(a_cov.to(dev))
        if init_actions is None:
            self.a_mat = torch.zeros(self.n_pol, *self.a_seq.shape, device=dev)
        else:
            assert init_actions.shape == (self.n_pol, *self.a_seq.shape), "Initial actions shape mismatch."
            self.a_mat = init_actions.clone().to(dev, torch.float32)
        self.a_mix = torch.ones(self.n_pol, device=dev)
        self._params_sampling = params_sampling
        self._params_log_space = params_log_space
        self._tf = None
        if params_sampling is False or params_sampling is None or (isinstance(params_sampling, str) and params_sampling == "none"):
            self.n_params, self._params_shape = 1, None
        elif params_sampling is True:
            self.n_params, self._params_shape = params_samples, [params_samples]
        elif isinstance(params_sampling, MerweScaledUTF):
            assert self._params_log_space is False, "Distribution must not be on log space if using UTF."
            if self.a_reg != 0:
                raise NotImplementedError("sigma-point rollouts with ctrl_penalty != 1: the reference forms that "
                                          "control cost from action sample 0 only (disco.py:334-344 on the "
                                          "unexpanded actions); not reproduced")
            self.n_params, self._params_shape = 1, None     # disco.py:128-131: sic
            self._tf = params_sampling
        else:
            raise ValueError("Invalid value for 'params_sampling': {}".format(params_sampling))
        self.n_rollouts = self.n_params * self.n_actions * self.n_pol
        self.return_states = return_states
        self._spec_cache = {}

    # ------------------------------------------------------------------------------------
    def _spec(self, model):
        key = id(model)
        fp = model.spec_fingerprint() if hasattr(model, "spec_fingerprint") else None
        ent = self._spec_cache.get(key)
        if ent is None or ent[2] != fp:     # first use, or the model's defaults / cost weights were changed in place
            ent = self._spec_cache[key] = (model, model.device_spec(self._inst_cost_fn, self._term_cost_fn, self.device), fp)
        return ent[1]

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, {} if k == "_spec_cache" else copy.deepcopy(v, memo))
        return new

    def _sample_params(self, model, params_dist):
        """disco.py:166-184: returns (device params [1,P,dp] | None, tiling, params_log_p)."""
        if self._params_shape is None:
            if hasattr(model, "defaults_are_tensors") and model.defaults_are_tensors():
                p = model._dict_to_device_params({}, 1, self.device)
                return p.reshape(1, 1, -1), L.PARAMS_BLOCKED, None
            return None, L.PARAMS_BLOCKED, None
        params = params_dist.sample(self._params_shape)
        drawn = params
        params_log_p = Lazy(lambda: params_dist.log_prob(drawn))     # read by diagnostics only: formed on first use
        if self._params_log_space is True:
            params = params.exp()
        tiling = L.PARAMS_INTERLEAVED if params.ndim == 1 else L.PARAMS_BLOCKED
        if params.ndim == 1:
            params = params.reshape(-1, 1)
        dev_params = model.device_params(params, self.device).unsqueeze(0)
        return dev_params, tiling, params_log_p

    def _sigma_points(self, model, params_dist):
        """disco.py:238-251, 262-264, 285-291: sigma points of the parameter belief as device params
        [1, pts, dp], their weights, and the (identical for every rollout) weighted log-probability."""
        try:
            cov, mean = params_dist.covariance_matrix, params_dist.mean
        except AttributeError:
            cov, mean = params_dist.variance.diag(), params_dist.mean
        params_sp = self._tf.compute_sigma_points(mean.cpu(), cov.cpu())          # [n, pts]
        pts = params_sp.T.contiguous()                                             # row k = sigma point k
        w = self._tf.loc_weights
        log_p = params_dist.log_prob(pts.to(mean.device))                          # [pts]
        log_p = log_p.expand(self.n_actions, self.n_pol, self._tf.pts) @ w.to(log_p.device)
        return model.device_params(pts, self.device).unsqueeze(0), w.to(self.device).contiguous(), log_p

    def evaluate(self, state, model, params_dist, noise, theta=None, want=("costs",), likelihood=L.LIK_EXP_UTILITY,
                 alpha=1.0, pert=None):
        """Run K1 for this controller's shapes.  noise [S,N,H,A] (eps when theta is given, else actions)."""
        spec = self._spec(model)
        dev = self.device
        state0 = torch.as_tensor(state, dtype=torch.float32).reshape(1, -1).to(dev).contiguous()
        sigma_w = None
        if self._tf is not None:
            params, sigma_w, params_log_p = self._sigma_points(model, params_dist)
            tiling = L.PARAMS_BLOCKED
        else:
            params, tiling, params_log_p = self._sample_params(model, params_dist)
        want = tuple(want) + (("states",) if self.return_states and "states" not in want else ())
        ctrl_mat = None
        if self.a_reg != 0:
            # the regulariser reads a_mat, and the reference's forward moves a_mat / a_mix on EVERY call, also
            # when SVMPC drives it with external actions (disco.py:392-393): keep both in step
            ctrl_mat = (self.a_mat @ self.a_pre).unsqueeze(0).contiguous()
            want = want + tuple(k for k in ("mppi_delta", "mix") if k not in want)
        out = ops.rollout_cost(
            spec, state0, noise.unsqueeze(0), theta=None if theta is None else theta.unsqueeze(0),
            sigma=self._sigma if theta is not None else None, params=params, param_tiling=tiling,
            likelihood=likelihood, a_seq=self.a_seq.unsqueeze(0).contiguous(),
            pert=None if pert is None else pert.unsqueeze(0), alpha=alpha, temperature=self.temp, want=want,
            sigma_weights=sigma_w, ctrl_mat=ctrl_mat, ctrl_reg=self.a_reg)
        res = {k: v[0] for k, v in out.items()}
        if self._tf is not None and "states" in res:
            # the reference's row order is (sample, policy, sigma point) relabelled as [S*pts, N] (disco.py:281-284)
            st = res["states"]                                   # [pts, S, N, H+1, ds]
            res["states"] = st.permute(1, 2, 0, 3, 4).reshape(self.n_actions * self._tf.pts, self.n_pol, self.hz_len + 1,
                                                              self.dim_s)
        res["params_log_p"] = params_log_p
        return res

    def forward(self, state, model, params_dist=None, ext_actions=None, debug=False):
        """disco.py:348-394 -> (costs [S,N], states [P,S,N,H+1,ds], actions [P,S,N,H,A], weights [S,N],
        params_log_p [P])."""
        dev = self.device
        pert = None
        if ext_actions is None:
            eps = self.a_dist.sample(sample_shape=[self.n_actions, self.n_pol, self.hz_len])
            actions = (eps + self.a_mat).contiguous()
            pert = eps.contiguous()
        else:
            actions = torch.as_tensor(ext_actions, dtype=torch.float32).to(dev).contiguous().clone()
        res = self.evaluate(state, model, params_dist, actions, want=("costs", "mppi_weights", "mppi_delta", "mix"),
                            pert=pert)
        self.a_mat += res["mppi_delta"]
        self.a_mix = res["mix"]
        states = res.get("states")
        # _sigma_rollout hands the actions back as given, _rollout tiled over the parameter draws
        acts = actions if self._tf is not None else actions.unsqueeze(0).expand(self.n_params, -1, -1, -1, -1)
        return res["costs"], states, acts, res["mppi_weights"], resolve(res["params_log_p"])

    def step(self, strategy="argmax", steps=1, ext_actions=None):
        """disco.py:396-417."""
        if strategy == "external" and ext_actions is not None:
            a_seq = torch.as_tensor(ext_actions, dtype=torch.float32).to(self.device).clone()
            a_seq = torch.max(torch.min(a_seq, self.max_a), self.min_a)
            nxt = a_seq[:steps].clone()
            self.a_seq = a_seq.roll(shifts=-steps, dims=0)
            self.a_seq[-steps:] = 0
            self.a_mat = self.a_mat.roll(shifts=-steps, dims=1)
            self.a_mat[:, -steps:] = 0
            return nxt
        if strategy not in ("argmax", "average"):
            raise ValueError("Invalid value for strategy.")
        a_mat = self.a_mat.unsqueeze(0).contiguous()
        nxt, a_seq = ops.disco_step(a_mat, self.a_mix.unsqueeze(0).contiguous(), self.min_a, self.max_a,
                                    L.SELECT_ARGMAX if strategy == "argmax" else L.SELECT_AVERAGE, steps)
        self.a_mat = a_mat[0]
        self.a_seq = a_seq[0]
        return nxt[0]
