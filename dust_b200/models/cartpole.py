"""Cart-pole with a continuous force in [-1, 1] * f_mag (API of dust/models/cartpole.py:8-172).  The reference's own
`step` raises AttributeError (it reads the name-mangled `self.__params_dict`, cartpole.py:150-155); this class evaluates
the method body as written -- `mass = m_c + m_c` (sic) included -- with that lookup repaired.  No cost function or demo
ships for it in the reference: only `step` exists, one thread per row on the GPU."""
import math

import torch

from .. import _lib as L
from .. import ops
from ..utils.spaces import Box
from .base import BaseModel

AUX_CARTPOLE = 1


class CartPoleModel(BaseModel):
    device_param_order = ("g", "mass_cart", "mass_pole", "length", "mu_c", "mu_p", "f_mag")

    def __init__(self, g=9.8, f_mag=10.0, mass_cart=1.0, mass_pole=0.1, length=1.0, mu_c=0.5e-3, mu_p=2e-6, **kwargs):
        params_dict = {"g": g, "mass_cart": mass_cart, "mass_pole": mass_pole, "length": length, "mu_c": mu_c, "mu_p": mu_p,
                       "f_mag": f_mag}
        super().__init__(params_dict=params_dict, **kwargs)
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        high = torch.tensor([self.x_threshold * 2, float("Inf"), self.theta_threshold_radians * 2, float("Inf")])
        self._action_space = Box(dim=1, low=-1, high=1, dtype=torch.float32)
        self._observation_space = Box(4, -high, high, dtype=torch.float32)

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def action_space(self):
        return self._action_space

    def step(self, states, actions, params_dict=None):
        """states [M,4] = (x, x_dot, theta, theta_dot), actions [M,1], params_dict {key: [M,1] | [1,1]} or None
        -> states + delta * dt (cartpole.py:127-172)."""
        L.require_cuda()
        dev = states.device if torch.is_tensor(states) and states.is_cuda else torch.device("cuda", torch.cuda.current_device())
        st = torch.as_tensor(states, dtype=torch.float32).reshape(-1, 4).to(dev).contiguous()
        M = st.shape[0]
        ac = torch.as_tensor(actions, dtype=torch.float32).reshape(-1, 1).to(dev).expand(M, 1).contiguous()
        prm = self._dict_to_device_params(params_dict, M, dev)
        cfg = [float(self.params_dict[k]) for k in self.device_param_order]
        return ops.aux_model_step(AUX_CARTPOLE, self.dt, cfg, st, ac, prm)
