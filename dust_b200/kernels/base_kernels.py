"""RBF kernel object (API of dust/kernels/base_kernels.py:39-108).  Inside SVMPC the kernel object
only selects which (gamma, c1, c2) the fused phi kernel uses; `eval` materialises K and dK for
small inputs with device tensor ops (diagnostic use, not the hot path)."""
import numpy as np
import torch


class BaseKernel:
    def __init__(self, analytic_grad=True):
        self.analytic_grad = analytic_grad


class RBF(BaseKernel):
    """k(x, x') = exp(-|x - x'|^2 / h), h = scale * median / log(N + 1) (median heuristic) or
    ell^2 / log(N + 1)."""

    def __init__(self, bandwidth=-1, bw_scale=1.0, analytic_grad=True, minimum_bw=1e-5, **kwargs):
        super().__init__(analytic_grad)
        self.ell = bandwidth
        self.ell_scale = bw_scale
        self.minimum_bw = minimum_bw

    def compute_bandwidth(self, X, Y):
        d2 = -2 * X.matmul(Y.t()) + (X * X).sum(-1).unsqueeze(1) + (Y * Y).sum(-1).unsqueeze(0)
        if self.ell < 0:
            h = torch.median(d2).detach()
        else:
            h = torch.as_tensor(self.ell ** 2, dtype=X.dtype, device=X.device)
        h = h / np.log(X.shape[0] + 1)
        h = torch.clamp(self.ell_scale * h, min=self.minimum_bw)
        return h, d2

    def eval(self, X, Y):
        assert X.shape == Y.shape
        if not self.analytic_grad:
            raise NotImplementedError
        h, d2 = self.compute_bandwidth(X, Y)
        K = (-d2 / h).exp()
        dK = K.unsqueeze(2) * (X.unsqueeze(1) - Y) * 2 / h
        return K, dK


class RBFKernel(torch.nn.Module):
    """Stand-in for `gpytorch.kernels.RBFKernel()` as the reference's demos construct it
    (default hyper-parameters: lengthscale = softplus(0) = ln 2, never changed -- the reference
    assigns a misspelt attribute, dust/inference/svmpc.py:78).  gpytorch is not a dependency of
    this package; SVMPC accepts this class or a real gpytorch RBFKernel and, either way, only
    reads `lengthscale`."""

    def __init__(self):
        super().__init__()
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(1, 1))

    @property
    def lengthscale(self):
        return torch.nn.functional.softplus(self.raw_lengthscale)
