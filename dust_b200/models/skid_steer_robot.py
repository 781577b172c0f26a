"""Skid-steer robot, simplified kinematic model (API of dust/models/skid_steer_robot.py:9-122).  The reference ships
no cost function and no demo for it: only `step` exists, here as one thread per row on the GPU."""
import torch

from .. import _lib as L
from .. import ops
from ..utils.spaces import Box
from .base import BaseModel

AUX_SKID_STEER = 0


class SkidSteerRobot(BaseModel):
    device_param_order = ("x_icr", "wheel_radius", "axial_distance")

    def __init__(self, delta_t, x_icr=0.2, wheel_radius=0.0625, axial_distance=0.475, min_wheel_speed=-0.5,
                 max_wheel_speed=0.5, **kwargs):
        params_dict = {"x_icr": x_icr, "wheel_radius": wheel_radius, "axial_distance": axial_distance}
        super().__init__(dt=delta_t, params_dict=params_dict, **kwargs)
        self._observation_space = Box(dim=5, low=-float("inf"), high=float("inf"), dtype=torch.float)
        self._action_space = Box(dim=2, low=min_wheel_speed, high=max_wheel_speed, dtype=torch.float)

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def action_space(self):
        return self._action_space

    def step(self, states, actions, params_dict=None):
        """states [M,5] = (x, y, theta, v, omega), actions [M,2] = (right, left) wheel speeds, params_dict {key: [M,1] | [1,1]}
        or None -> next states [M,5] (skid_steer_robot.py:73-122)."""
        L.require_cuda()
        dev = states.device if torch.is_tensor(states) and states.is_cuda else torch.device("cuda", torch.cuda.current_device())
        st = torch.as_tensor(states, dtype=torch.float32).reshape(-1, 5).to(dev).contiguous()
        M = st.shape[0]
        ac = torch.as_tensor(actions, dtype=torch.float32).reshape(-1, 2).to(dev).expand(M, 2).contiguous()
        prm = self._dict_to_device_params(params_dict, M, dev)
        lo, hi = self.action_space.low, self.action_space.high
        cfg = [float(self.params_dict[k]) for k in self.device_param_order] + [float(lo[0]), float(hi[0]), float(lo[1]), float(hi[1])]
        return ops.aux_model_step(AUX_SKID_STEER, self.dt, cfg, st, ac, prm)
