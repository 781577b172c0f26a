"""Generic SVGD (API of dust/inference/svgd.py:28-187) with the kernel work on the GPU: exact
median bandwidth by radix select (K4) and the fused phi kernel (K5)."""
import torch
import torch.distributions as dist
import torch.optim as optim

from .. import ops


def squared_distance(x1, x2):
    """Materialised [N,M] squared distances (svgd.py:28-39); small inputs / diagnostics only."""
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
    res = torch.addmm(x2_norm.transpose(-2, -1), x1, x2.transpose(-2, -1), alpha=-2).add_(x1_norm)
    return res.clamp(min=0)


def bw_median(x, y=None, bw_scale=1.0, tol=1.0e-5):
    """svgd.py:42-52: bw = scale * max(sqrt(median/2) / log(N+1), tol), with the exact lower
    median of all N^2 squared distances found on the device without materialising them."""
    if y is not None and y is not x and not torch.equal(x, y):
        raise NotImplementedError("bw_median: only the x == y case has a device kernel")
    xd = x.detach().to("cuda", torch.float32).contiguous()
    med = ops.median_sq_dist(xd)[0]
    h = torch.sqrt(0.5 * med) / torch.tensor(x.shape[0] + 1.0).log().to(med.device)
    return bw_scale * h.clamp_min(tol)


def get_gmm(x, weights, covariance):
    mix = dist.Categorical(weights)
    comp = dist.Independent(dist.MultivariateNormal(x.detach(), covariance), 1)
    return dist.mixture_same_family.MixtureSameFamily(mix, comp)


def default_kernel(x, y=None, bw=0.69):
    if y is None:
        y = x.detach().clone().flatten(1, -1)
    return torch.exp(-squared_distance(x, y) / bw ** 2 / 2)


class SVGD:
    def __init__(self, kernel=None, bw_scale=1.0, n_particles=None, n_steps=100, optimizer_class=optim.Adam,
                 **opt_args):
        self.kernel = default_kernel if kernel is None else kernel
        self.bw_scale = bw_scale
        self.optimizer_class = optimizer_class
        self.opt_args = opt_args
        self.n_steps = n_steps
        self.n_particles = n_particles

    def phi(self, x, log_p, h):
        """svgd.py:127-135: (K grad log p + sum grad k) / N with K = exp(-d2 / (2 h^2)).  The score
        comes from the caller's log_p (torch autograd on the device); K is never materialised."""
        if self.kernel is not default_kernel:
            raise NotImplementedError("SVGD.phi: only the default RBF kernel has a device kernel")
        xg = x.detach().to("cuda", torch.float32).requires_grad_(True)
        score = torch.autograd.grad(log_p(xg).sum(), xg)[0]
        return self.phi_from_score(xg.detach(), score, h)

    @staticmethod
    def phi_from_score(x, score, h):
        N = x.shape[0]
        h = float(h)
        xf = x.reshape(1, N, -1).contiguous()
        sf = score.reshape(1, N, -1).contiguous()
        out = ops.svgd_phi(xf, sf, gamma=1.0 / (2.0 * h * h), c1=1.0 / N, c2=1.0 / (N * h * h))
        return out["phi"].reshape(x.shape)

    def step(self, x, optimizer, log_p, bw):
        optimizer.zero_grad()
        x.grad = -self.phi(x, log_p, bw)
        optimizer.step()

    def optimize(self, log_p, initial_particles=None, prior=None, debug=False, bw=0.69):
        if initial_particles is not None:
            x = initial_particles.detach().clone().to("cuda", torch.float32).requires_grad_(True)
        elif prior is not None:
            x = prior.sample(torch.Size([self.n_particles])).to("cuda", torch.float32).requires_grad_(True)
        else:
            raise RuntimeError("Either initial_particles or prior must be specified for SVGD")
        optimizer = self.optimizer_class(params=[x], **self.opt_args)
        if self.kernel is default_kernel:
            bw = float(bw_median(x, x))
        for _ in range(self.n_steps):
            self.step(x, optimizer, log_p, bw)
        return x.detach()
