"""Accuracy of one build of the library (DUST_B200_LIB=...) on the pendulum fixtures: element-wise cost
error against the golden vectors and phi error against golden / float64.  Used to decide which
instruction-count reductions of the pendulum step are kept (DUST_PEND_OPT, models.cuh)."""
import os, sys, time
import torch
sys.path.insert(0, os.getcwd())
from oracle import dust_oracle as O
from tests.util import load, rel_elem, rel_max, golden_grid
from tests.test_gpu_parity import cu, dev_params, HYPER
from dust_b200 import ops, _lib as L
from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost
from dust_b200.inference.core import SvmpcCore

pend = PendulumModel(uncertain_params=("length", "mass"))
spec = pend.device_spec(inst_cost, term_cost, "cuda")
model = O.Model("pendulum", O.ParticleCfg(golden_grid()))
tag = os.environ.get("DUST_B200_LIB", "default")
res = []
print("==", tag, flush=True)
for seed in (0, 1, 2):
    d = load(f"fwd_pendulum_s{seed}")
    params, tiling = dev_params("pendulum", d["params"], bool(d["log_space"]))
    out = ops.rollout_cost(spec, cu(d["state"]).reshape(1, -1), cu(d["actions"]).unsqueeze(0), params=params,
                           param_tiling=tiling, temperature=float(d["temp"]), want=("costs",))
    print(f"fwd_s{seed} cost {rel_elem(out['costs'][0].cpu(), d['costs']):.2e}")
for name, kern in (("svmpc_pendulum_rbf", "gpytorch"), ("svmpc_pendulum_mp", "mp"), ("dual_pendulum_bw", "gpytorch")):
    d = load(name)
    c = HYPER["pendulum"]
    for t in range(int(d["n_steps"])):
        gi, go = (lambda k: d[f"t{t}_in_{k}"]), (lambda k: d[f"t{t}_out_{k}"])
        params, tiling = dev_params("pendulum", gi("params"), c["log"])
        core = SvmpcCore(spec, cu(gi("theta0")).unsqueeze(0), cu(gi("mu0")).unsqueeze(0), cu(gi("mix0")).unsqueeze(0),
                         torch.full((1,), c["var"]), d["sigma"], alpha=c["alpha"], temperature=1.0 / c["alpha"], lr=c["lr"],
                         kernel=kern, weighted_prior=c["wp"], aliased=t > 0)
        out = core.optimize_step(cu(gi("state")).reshape(1, -1), cu(gi("eps")).unsqueeze(0), params, tiling)
        st = O.SvmpcState(gi("theta0").double(), gi("mu0").double(), gi("mix0").double(), c["var"], aliased=t > 0)
        p64 = gi("params").double() if gi("params").numel() else None
        ref = O.svmpc_optimize(model, st, gi("state").double(), gi("eps").double(), d["sigma"].double(), p64, c["log"],
                               c["alpha"], c["lr"], kernel="rbf" if kern == "gpytorch" else "mp")
        costs = out["costs"][0].cpu()
        print(f"{name} t{t}: cost {rel_elem(costs, go('costs')):.1e} abs {float((costs-go('costs')).abs().max()):.1e} "
                   f"cmax {float(go('costs').max()):.0f} | phi gold {rel_max(out['phi'][0].cpu(), go('phi')):.1e} "
                   f"f64 {rel_max(out['phi'][0].cpu(), ref['phi']):.1e} refnoise {rel_max(go('phi'), ref['phi']):.1e}")


