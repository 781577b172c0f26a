"""Message-passing kernel object (API of dust/kernels/composite_kernels.py:9-64): an independent
scalar RBF per (timestep, control) column, each with its own median bandwidth.  SVMPC maps it
onto the per-dimension mode of the fused phi kernel."""
from .base_kernels import RBF


class CompositeKernel:
    def __init__(self, base_kernel=None, ctrl_dim=1, indep_controls=True, **kwargs):
        self.ctrl_dim = ctrl_dim
        self.base_kernel = RBF() if base_kernel is None else base_kernel
        self.indep_controls = indep_controls


class iid_mp(CompositeKernel):
    def eval(self, X, Y, **kwargs):
        import torch

        if not self.indep_controls:
            raise NotImplementedError("iid_mp(indep_controls=False) has no device kernel")
        m, D = X.shape[0], X.reshape(X.shape[0], -1).shape[1]
        Xf, Yf = X.reshape(m, D), Y.reshape(m, D)
        K = torch.zeros(m, m, D, dtype=X.dtype, device=X.device)
        dK = torch.zeros(m, m, D, dtype=X.dtype, device=X.device)
        for q in range(D):
            k, dk = self.base_kernel.eval(Xf[:, q:q + 1], Yf[:, q:q + 1])
            K[:, :, q], dK[:, :, q] = k, dk.squeeze(2)
        return K, dK
