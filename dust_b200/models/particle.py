"""2-D point mass among obstacles (dust/models/particle.py:117-225) on the GPU."""
import torch

from .. import _lib as L
from .. import ops
from ..utils.obstacle_map import generate_obstacle_map, get_obst_preset
from ..utils.spaces import Box
from .base import BaseModel, DeviceModelSpec


class Particle(BaseModel):
    device_param_order = ("mass",)

    def __init__(self, mass=1.0, noise_std=torch.zeros(2), control_type="acceleration", cost_params=None,
                 with_obstacle=False, obst_preset=None, obst_width=None, obst_params=None, map_size=None,
                 map_type=None, map_cell_size=None, init_state=None, target_state=None, can_crash=False,
                 max_speed=None, max_accel=None, verbose=False, deterministic=False, euler_steps=1, **kwargs):
        super().__init__(params_dict={"mass": mass}, **kwargs)
        if control_type != "acceleration":
            if control_type == "velocity":
                raise NotImplementedError("Particle: control_type='velocity' has no device kernel")
            raise IOError('control_type "{}" not recognized'.format(control_type))
        if not deterministic:
            raise NotImplementedError("Particle: control-channel noise (deterministic=False) has no device kernel")
        self._max_speed = float("inf") if max_speed is None else float(max_speed)
        self._max_acc = float("inf") if max_accel is None else float(max_accel)
        bounds = torch.tensor([float("inf"), float("inf"), self._max_speed, self._max_speed])
        self._observation_space = Box(dim=4, low=-bounds, high=bounds, dtype=torch.float)
        self._action_space = Box(dim=2, low=-self._max_acc, high=self._max_acc, dtype=torch.float)
        self.target = torch.zeros(4) if target_state is None else torch.as_tensor(target_state, dtype=torch.float)
        self.dyn_std = noise_std
        self.init_state = torch.as_tensor(init_state)
        self.euler_steps = euler_steps
        self.control_type = control_type
        self.with_obstacle, self.can_crash = with_obstacle, can_crash
        self.map_cell_size, self.map_size = map_cell_size, map_size
        assert map_size[0] % 2 == 0 and map_size[1] % 2 == 0
        self.cmap_size = [torch.as_tensor(map_size[0] / map_cell_size).ceil(),
                          torch.as_tensor(map_size[1] / map_cell_size).ceil()]
        self.c_offset = torch.Tensor([int(self.cmap_size[0] / 2), int(self.cmap_size[1] / 2)])
        self.verbose, self.deterministic = verbose, deterministic
        self.init_cost_weights(cost_params)
        self.obst_params = obst_params
        self.obst_map = None
        if with_obstacle:
            self.obst_params = get_obst_preset(obst_preset, obst_width)
            self.obst_map = generate_obstacle_map(map_size, self.obst_params, map_cell_size, map_type=map_type)

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def action_space(self):
        return self._action_space

    def init_cost_weights(self, params):
        """dust/models/particle.py:292-326."""
        if params is None:
            params = dict.fromkeys(["w_qpos", "w_qvel", "w_qpos_T", "w_qvel_T", "w_ctrl", "w_obs"], 1.0)
        self.w_state = torch.as_tensor([params["w_qpos"]] * 2 + [params["w_qvel"]] * 2, dtype=torch.float)
        self.w_ctrl = torch.as_tensor([params["w_ctrl"]] * 2, dtype=torch.float)
        self.w_term = torch.as_tensor([params["w_qpos_T"]] * 2 + [params["w_qvel_T"]] * 2, dtype=torch.float)
        self.w_obs = torch.as_tensor([params["w_obs"]], dtype=torch.float)

    def device_spec(self, inst_cost_fn=None, term_cost_fn=None, device="cuda"):
        for fn, ok in ((inst_cost_fn, self.default_inst_cost), (term_cost_fn, self.default_term_cost)):
            if fn is None:
                continue
            same = getattr(fn, "__func__", None) is ok.__func__ and isinstance(getattr(fn, "__self__", None), Particle)
            if not same:
                raise NotImplementedError(
                    "Particle: only Particle.default_inst_cost / default_term_cost have a device kernel; got %r" % (fn,))
        d = L.ModelDesc()
        d.kind, d.dt = L.MODEL_PARTICLE, float(self.dt)
        d.default_mass = float(self.params_dict["mass"])
        big = 3.0e38
        d.max_accel, d.max_speed = min(self._max_acc, big), min(self._max_speed, big)
        for i in range(4):
            d.target[i], d.w_state[i], d.w_term[i] = float(self.target[i]), float(self.w_state[i]), float(self.w_term[i])
        d.w_ctrl[0], d.w_ctrl[1] = float(self.w_ctrl[0]), float(self.w_ctrl[1])
        d.w_obs = float(self.w_obs[0])
        keep = []
        if self.with_obstacle:
            bits = self.obst_map.device_bits(torch.device(device))
            keep.append(bits)
            d.inv_cell = float(1 / self.obst_map.cell_size)
            d.c_offset[0], d.c_offset[1] = float(self.obst_map.c_offset[0]), float(self.obst_map.c_offset[1])
            d.grid_nx, d.grid_ny = self.obst_map.map.shape
            d.grid_bits = bits.data_ptr()
        d.with_obstacle = int(bool(self.with_obstacle))
        d.can_crash = int(bool(self.can_crash and self.with_obstacle))
        return DeviceModelSpec(d, L.MODEL_PARTICLE, keep)

    def step(self, states, actions, params_dict=None):
        L.require_cuda()
        dev = states.device if torch.is_tensor(states) and states.is_cuda else torch.device("cuda", torch.cuda.current_device())
        shape = torch.as_tensor(states).shape
        st = torch.as_tensor(states, dtype=torch.float32).reshape(-1, 4).to(dev).contiguous()
        M = st.shape[0]
        ac = torch.as_tensor(actions, dtype=torch.float32).reshape(-1, 2).to(dev).expand(M, 2).contiguous()
        prm = self._dict_to_device_params(params_dict, M, dev)
        return ops.model_step(self.cached_spec(dev), st, ac, prm).reshape(shape)

    def _cost(self, states, actions, terminal):
        L.require_cuda()
        dev = states.device if torch.is_tensor(states) and states.is_cuda else torch.device("cuda", torch.cuda.current_device())
        st = torch.as_tensor(states, dtype=torch.float32).reshape(-1, 4).to(dev).contiguous()
        ac = None
        if actions is not None and torch.is_tensor(actions):
            ac = actions.to(dev, torch.float32).reshape(-1, 2).expand(st.shape[0], 2).contiguous()
        return ops.model_cost(self.cached_spec(dev), st, ac, terminal)

    def default_inst_cost(self, states, actions=0, n_pol=0, debug=False):
        return self._cost(states, actions, False)

    def default_term_cost(self, states, n_pol=0, debug=False):
        return self._cost(states, None, True)

    def to_map_coord(self, coord_vec):
        assert coord_vec.shape[-1] == 2, "Coordinates must be 2-D."
        return self.c_offset.to(coord_vec.device) + coord_vec / self.map_cell_size
