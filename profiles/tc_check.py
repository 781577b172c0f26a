import os, sys, torch
sys.path.insert(0, os.getcwd())
from dust_b200 import ops
from oracle import dust_oracle as O
torch.manual_seed(0)
for (N, D, gamma) in ((1024, 40, 0.02), (2048, 40, 0.02), (4096, 16, 0.05), (2048, 56, 0.01), (8192, 40, 0.012)):
    X = torch.randn(N, D); S = torch.randn(N, D)
    x, s = X.cuda().unsqueeze(0), S.cuda().unsqueeze(0)
    c1, c2 = 1.0 / N, 0.37 / N
    os.environ.pop("DUST_B200_NO_TC", None)
    tc = ops.svgd_phi(x, s, gamma=gamma, c1=c1, c2=c2)["phi"][0].cpu()
    torch.cuda.synchronize()
    os.environ["DUST_B200_NO_TC"] = "1"
    simt = ops.svgd_phi(x, s, gamma=gamma, c1=c1, c2=c2)["phi"][0].cpu()
    ref = O.phi_unified_tiled(X, S, gamma, c1, c2, tile=1024)
    e = lambda a: float((a.double() - ref).abs().max() / ref.abs().max())
    print(N, D, "tc err", e(tc), "simt err", e(simt), "tc-vs-simt", float((tc - simt).abs().max() / simt.abs().max()), flush=True)
