"""Forward-model protocol (dust/models/base.py:20-58,173-183) + the device description that the
CUDA kernels consume.  Only models with a hand-written device kernel can be used: anything
else raises NotImplementedError -- there is no traced / CPU fallback."""
import torch

from .. import _lib as L


class DeviceModelSpec:
    """A `dust_model_desc` (ctypes) plus the tensors it points to, kept alive together."""

    def __init__(self, desc, kind, keepalive=()):
        self.desc, self.kind = desc, kind
        self.ds = 2 if kind == L.MODEL_PENDULUM else 4
        self.da = 1 if kind == L.MODEL_PENDULUM else 2
        self.dp = 2 if kind == L.MODEL_PENDULUM else 1
        self._keepalive = tuple(keepalive)

    def __deepcopy__(self, memo):
        # an immutable description (ctypes struct + the device tensors it points to): copies of the objects
        # that hold it (the demos deep-copy controller / SVMPC / MPF per episode) share it
        return self


class BaseModel:
    def __init__(self, dt=0.05, params_dict=None, uncertain_params=None):
        assert dt > 0, "Delta t must be greater than zero."
        self._dt = dt
        self._params_dict = {} if params_dict is None else params_dict
        self._params_keys = uncertain_params

    @property
    def dt(self):
        return self._dt

    @property
    def params_dict(self):
        return self._params_dict

    @params_dict.setter
    def params_dict(self, params_dict):
        self._params_dict = params_dict

    @property
    def uncertain_params(self):
        return self._params_keys

    def set_params_from_dist(self, params_dist):
        for idx, key in enumerate(self.uncertain_params):
            self._params_dict[key] = params_dist.mean[idx]

    def params_to_dict(self, params):
        return {key: params[:, idx].reshape(-1, 1) for idx, key in enumerate(self._params_keys)}

    def dict_to_params(self, params_dict):
        return torch.cat([params_dict[key] for key in self._params_keys], dim=1)

    def spec_fingerprint(self):
        """What a DeviceModelSpec bakes in (default parameters, cost weights): holders of a cached spec compare it
        and rebuild the spec when the model was changed in place (the reference reads params_dict on every step)."""
        out = []
        for k in sorted(self._params_dict):
            v = self._params_dict[k]
            if torch.is_tensor(v):
                out.append((k, id(v), v._version) if v.is_cuda else (k, tuple(v.reshape(-1).tolist())))
            else:
                out.append((k, v))
        for name in ("w_state", "w_ctrl", "w_term", "w_obs", "target"):
            t = getattr(self, name, None)
            if torch.is_tensor(t):
                out.append((name, tuple(t.reshape(-1).tolist())))
        return tuple(out)

    def cached_spec(self, device):
        """The cost-free DeviceModelSpec of this model on `device` (what `step` needs), rebuilt only when the model's
        defaults changed."""
        fp = (str(device), self.spec_fingerprint())
        ent = self.__dict__.get("_step_spec")
        if ent is None or ent[0] != fp:
            ent = self.__dict__["_step_spec"] = (fp, self.device_spec(device=device))
        return ent[1]

    # --- device side -------------------------------------------------------------------
    #: order of the parameter columns the kernels expect
    device_param_order = ()

    def device_spec(self, inst_cost_fn, term_cost_fn, device):
        raise NotImplementedError(
            f"{type(self).__name__} has no device kernel; models with CUDA kernels: "
            "dust_b200.models.pendulum.PendulumModel, dust_b200.models.particle.Particle")

    def device_params(self, params, device):
        """Sampled parameters [P, len(uncertain_params)] -> kernel layout [P, dp] (float32,
        on `device`), filling the columns that are not sampled with the model defaults."""
        cols = []
        keys = list(self._params_keys) if self._params_keys else []
        for name in self.device_param_order:
            if name in keys:
                cols.append(params[:, keys.index(name)].reshape(-1, 1).to(torch.float32))
            else:
                cols.append(torch.full((params.shape[0], 1), float(self._params_dict[name]), dtype=torch.float32,
                                       device=params.device))
        return torch.cat(cols, dim=1).to(device).contiguous()

    def _dict_to_device_params(self, params_dict, M, device):
        """params_dict {key: [M,1] | [1,1]} (BaseModel.step contract) -> [M, dp] or None."""
        if params_dict is None:
            return None
        cols = []
        for name in self.device_param_order:
            v = params_dict.get(name, self._params_dict[name])
            v = torch.as_tensor(v, dtype=torch.float32).reshape(-1, 1).to(device)
            cols.append(v.expand(M, 1))
        return torch.cat(cols, dim=1).contiguous()
