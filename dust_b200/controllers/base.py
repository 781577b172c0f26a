"""Controller base (contract of dust/controllers/base.py:7-81: dims, bounds, `a_seq`, cost-function
plumbing, `roll`).  The autograd Jacobian/Hessian helpers of the reference are unused by the
hot path and are not provided."""
import torch


class BaseController:
    def __init__(self, observation_space, action_space, hz_len, inst_cost_fn=None, term_cost_fn=None,
                 init_actions=None, device="cuda"):
        self.device = torch.device(device)
        self.hz_len = hz_len
        self.dim_s = observation_space.dim
        self.dim_a = action_space.dim
        self.min_a = action_space.low.to(self.device)
        self.max_a = action_space.high.to(self.device)
        if init_actions is None:
            self.a_seq = torch.zeros((self.hz_len, self.dim_a), device=self.device)
        else:
            self.a_seq = torch.as_tensor(init_actions, dtype=torch.float32).to(self.device)
        if inst_cost_fn is None and term_cost_fn is None:
            raise ValueError("Specify at least one cost function")
        self._inst_cost_fn = inst_cost_fn
        self._term_cost_fn = term_cost_fn

    @staticmethod
    def _null_cost_fn(state, *args, **kwargs):
        return (state * 0).sum(-1)

    @property
    def inst_cost_fn(self):
        return self._null_cost_fn if self._inst_cost_fn is None else self._inst_cost_fn

    @property
    def term_cost_fn(self):
        return self._null_cost_fn if self._term_cost_fn is None else self._term_cost_fn

    def roll(self, steps=1):
        self.a_seq = torch.roll(self.a_seq, -steps, 0)
        self.a_seq[-steps:, :] = 0

    def forward(self, model, state):
        raise NotImplementedError("Should be implemented by the subclass")
