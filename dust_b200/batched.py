"""Batched SVMPC: B independent MPC instances advanced by the same kernel launches (the
"4096 instances x 8 policies x 256 samples" configuration).  New API (the reference has no
batch dimension); same per-instance semantics as `dust_b200.inference.svmpc.SVMPC`."""
import torch

from . import _lib as L
from . import ops
from .inference.core import GPYTORCH_DEFAULT_LENGTHSCALE, SvmpcCore


class BatchedSVMPC:
    def __init__(self, model, n_instances, n_policies, action_samples, horizon, ctrl_sigma, prior_sigma,
                 alpha=1.0, learning_rate=1.0, kernel="gpytorch", inst_cost_fn=None, term_cost_fn=None,
                 weighted_prior=False, roll_strategy="repeat", grad="analytic", params_samples=0,
                 device="cuda", seed=0, noise_stream=0):
        L.require_cuda()
        self.device = torch.device(device)
        self.model = model
        self.spec = model.device_spec(inst_cost_fn, term_cost_fn, self.device)
        B, N, H, A = n_instances, n_policies, horizon, self.spec.da
        self.B, self.N, self.S, self.H, self.A, self.P = B, N, action_samples, H, A, params_samples
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        mu = torch.randn(B, N, H, A, device=self.device, generator=self.gen)
        # theta0 ~ GMM(mu, prior_sigma^2) with uniform weights (demo/pendulum_example.py:66-71)
        idx = torch.randint(0, N, (B, N), device=self.device, generator=self.gen)
        theta = torch.gather(mu, 1, idx[:, :, None, None].expand(B, N, H, A)) + prior_sigma * torch.randn(
            B, N, H, A, device=self.device, generator=self.gen)
        mix = torch.ones(B, N, device=self.device)
        sigma = torch.full((A,), float(ctrl_sigma))
        self.core = SvmpcCore(self.spec, theta, mu, mix, torch.full((A,), float(prior_sigma) ** 2), sigma,
                              alpha=alpha, temperature=1.0 / alpha, lr=learning_rate, kernel=kernel,
                              lengthscale=GPYTORCH_DEFAULT_LENGTHSCALE, grad=grad, roll_strategy=roll_strategy,
                              weighted_prior=weighted_prior)
        self.eps = torch.empty(B, self.S, N, H, A, device=self.device)
        # action noise: the library's counter-based generator, one Philox stream per (rank, draw)
        self.seed, self.noise_stream, self.draws = int(seed), int(noise_stream), 0

    @property
    def theta(self):
        return self.core.theta

    def draw_noise(self):
        ops.noise_normal(self.eps, self.seed, (self.noise_stream << 40) + self.draws)
        self.draws += 1
        return self.eps

    def optimize(self, state, eps=None, params=None):
        """state [B,ds] (device), eps [B,S,N,H,A] standard normal (drawn on the device if None)."""
        eps = self.draw_noise() if eps is None else eps
        return self.core.optimize_step(state, eps, params)

    def forward(self):
        """-> (a_seq [B,H,A], p_weights [B,N], i_star [B])"""
        return self.core.forward_step()

    def control_step(self, state, eps=None, params=None):
        """optimize + forward for every instance; returns the actions to apply [B,A]."""
        eps = self.draw_noise() if eps is None else eps
        a_seq, _, _ = self.core.control_step(state, eps, params)
        return a_seq[:, 0]
