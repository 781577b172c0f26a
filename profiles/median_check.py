import os, sys, torch
sys.path.insert(0, os.getcwd())
from dust_b200 import ops, _lib as L
torch.manual_seed(0)
lib = L.load()
def run(name, X):
    x = X.cuda().contiguous()
    ws = ops.MedianWorkspace(x.shape[0], x.shape[1], x.device)
    fast = float(ops.median_sq_dist(x, ws=ws)[0]); ok = int(ws.selected[5])
    robust = float(ops.median_sq_dist(x, allow_fast=False)[0])
    print(f"{name:28s} N={x.shape[0]:6d} D={x.shape[1]:3d} fast={fast:.9g} ok_flag={ok} robust={robust:.9g} rel_diff={abs(fast-robust)/max(abs(robust),1e-30):.2e}", flush=True)
run("normal", torch.randn(2048, 40))
run("normal", torch.randn(8192, 40))
run("anisotropic", torch.randn(4096, 40) * (torch.arange(1, 41).float() / 10))
run("low-dim", torch.randn(4096, 8))
run("heavy-tailed (cauchy)", torch.distributions.Cauchy(0., 1.).sample((4096, 16)))
run("two clusters", torch.cat([torch.randn(2048, 24), torch.randn(2048, 24) + 30.0]))
run("all identical", torch.ones(2048, 40) * 0.3)
run("half duplicates", torch.cat([torch.zeros(3072, 16), torch.randn(1024, 16)]))
