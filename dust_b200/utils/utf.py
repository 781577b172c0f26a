"""Merwe scaled sigma points for the unscented transform (API of dust/utils/utf.py:4-141).

Set-up-time host code: 2n+1 weights and an n x n matrix square root per control step.  The rollouts
of the sigma points and the weighted costs run on the device (`dust_rollout_args.sigma_weights`)."""
import torch


def _upper_cholesky(A):
    """U with U^T U = A (the reference's default `sqrt`, utf.py:45-50)."""
    return torch.linalg.cholesky(A.transpose(-2, -1).conj()).transpose(-2, -1).conj()


class MerweScaledUTF:
    def __init__(self, n, alpha=1e-3, beta=2, kappa=0, sqrt_method=None):
        self.n = n
        self.pts = 2 * n + 1
        self.alpha, self.beta, self.kappa = alpha, beta, kappa
        self.sqrt = _upper_cholesky if sqrt_method is None else sqrt_method
        self._compute_weights()

    @property
    def _lambda(self):
        return self.alpha ** 2 * (self.n + self.kappa) - self.n

    @property
    def loc_weights(self):
        """Weights of the mean, float32 [2n+1] (utf.py:81-91)."""
        return self._loc_weights

    @property
    def cov_weights(self):
        """Weights of the covariance, float32 [2n+1]."""
        return self._cov_weights

    def _compute_weights(self):
        n, lam = self.n, self._lambda
        c = 0.5 / (n + lam)
        self._loc_weights = torch.full((self.pts,), c, dtype=torch.float)
        self._cov_weights = torch.full((self.pts,), c, dtype=torch.float)
        self._loc_weights[0] = lam / (n + lam)
        self._cov_weights[0] = lam / (n + lam) + (1 - self.alpha ** 2 + self.beta)

    def compute_sigma_points(self, mu, K):
        """mu [n], K [n,n] -> sigmas [n, 2n+1]: column 0 the mean, then mean + rows... (utf.py:93-123:
        `U + mu.view(-1, 1)` adds mu[i] to ROW i of the upper factor, so column k of the block is
        mu + U[:, k])."""
        mu = torch.as_tensor(mu, dtype=torch.float)
        K = torch.as_tensor(K, dtype=torch.float)
        if self.n != mu.size(0):
            raise ValueError("expected size(x) {}, but size is {}".format(self.n, mu.size(0)))
        n = self.n
        U = self.sqrt((self._lambda + n) * K)
        col = mu.view(-1, 1)
        return torch.cat([col, U + col, -U + col], dim=1)

    def unscented_transform(self, sigmas):
        """sigmas [n, 2n+1] -> (mean [n], covariance [n,n]) (utf.py:125-141)."""
        mu = sigmas @ self._loc_weights
        y = sigmas - mu.view(-1, 1)
        return mu, y @ torch.diag(self._cov_weights) @ y.t()

    def __repr__(self):
        return "{}(n: {}, alpha: {}, beta: {}, kappa: {},\nloc_weights:\n{},\ncov_weights:\n{})".format(
            type(self).__name__, self.n, self.alpha, self.beta, self.kappa, self._loc_weights, self._cov_weights)
