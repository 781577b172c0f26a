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
                 device="cuda", seed=0, noise_stream=0, prefetch_noise=False):
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
        # prefetch (control_step): the draw of the NEXT step is enqueued on a side stream while this step's kernel runs --
        # it depends on nothing but the draw counter.  Two buffers alternate; events order fill -> use -> refill.
        self.prefetch = bool(prefetch_noise)
        self._side = torch.cuda.Stream(device=self.device) if self.prefetch else None   # a higher stream priority changes nothing (measured)
        self._eps2 = [self.eps, torch.empty_like(self.eps)] if self.prefetch else None
        self._filled = [None, None]        # event: buffer i holds the draw for the step that will use it
        self._used = [None, None]          # event: the step that read buffer i has been enqueued and finished
        self._next = 0

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

    def _enqueue_fill(self, i):
        """fill buffer i with the next draw on the side stream, after the step that last read it"""
        with torch.cuda.stream(self._side):
            if self._used[i] is not None:
                self._side.wait_event(self._used[i])
            ops.noise_normal(self._eps2[i], self.seed, (self.noise_stream << 40) + self.draws)
            self.draws += 1
            ev = torch.cuda.Event()
            ev.record(self._side)
            self._filled[i] = ev

    def control_step(self, state, eps=None, params=None):
        """optimize + forward for every instance; returns the actions to apply [B,A].
        With `prefetch_noise=True` (and no `eps` given) the same draws are made in the same order, one step ahead, on a
        side stream: the fill of step k + 1 shares the GPU with the kernel of step k instead of preceding it."""
        if eps is None and self.prefetch:
            cur = torch.cuda.current_stream(self.device)
            i = self._next
            if self._filled[i] is None:
                self._enqueue_fill(i)                      # first call: nothing was prefetched yet
            cur.wait_event(self._filled[i])
            self._filled[i] = None
            self._enqueue_fill(1 - i)                      # the next step's draw, concurrent with this step's kernel
            a_seq, _, _ = self.core.control_step(state, self._eps2[i], params)
            ev = torch.cuda.Event()
            ev.record(cur)
            self._used[i] = ev
            self._next = 1 - i
            return a_seq[:, 0]
        eps = self.draw_noise() if eps is None else eps
        a_seq, _, _ = self.core.control_step(state, eps, params)
        return a_seq[:, 0]
