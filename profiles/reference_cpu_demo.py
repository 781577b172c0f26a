"""Time the UNMODIFIED reference (lubaroli/dust, imported from /root/reference through oracle/refshim.py) on the two
demo configurations, in the build container (the GPU box has no /root/reference): one dual control step =
svmpc.optimize + svmpc.forward + plant step + mpf.optimize(20 steps), as demo/particle_example.py:177-207 runs it.
Anomaly mode is left as the reference sets it (dust/inference/svgd.py:7).  Writes profiles/r2_reference_cpu_demo.json.

usage: python profiles/reference_cpu_demo.py"""
import importlib.util
import json
import os
import platform
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)


def build(kind):
    if kind == "pendulum":
        w = G.make_pendulum(2, kernel="rbf")
        ep = w["cfg"]
        x0 = w["dyn"].sample([ep["mpf_n_particles"]])
        lik = G.likelihoods.GaussianLikelihood(initial_obs=w["state"].clone(), obs_std=ep["mpf_obs_std"],
                                               model=G.pendulum_mod.PendulumModel(uncertain_params=("length", "mass")), log_space=False)
        mpf = G.mpf_mod.MPF(init_particles=x0, likelihood=lik, optimizer_class=torch.optim.SGD, lr=ep["mpf_learning_rate"],
                            bw=None, bw_scale=ep["mpf_bandwidth_scaling"])
        pp = {"length": torch.tensor([[1.1]]), "mass": torch.tensor([[0.8]])}
        bw = ep["mpf_bandwidth"]
    else:
        w = G.make_particle(2, kernel="rbf")
        ep = w["cfg"]
        x0 = w["dyn"].sample([ep["mpf_n_particles"], 1]).clamp(min=1e-6).log()
        lik = G.likelihoods.GaussianLikelihood(initial_obs=w["state"].clone(), obs_std=ep["mpf_obs_std"], model=w["model"], log_space=True)
        mpf = G.mpf_mod.MPF(init_particles=x0, likelihood=lik, optimizer_class=torch.optim.SGD, lr=ep["mpf_learning_rate"],
                            bw=(2 * ep["dyn_prior_arg2"]) ** 1 / 2, bw_scale=ep["mpf_bandwidth_scaling"])
        pp = {"mass": torch.tensor([[3.0]])}
        bw = ep["mpf_bandwidth"]
    return w, mpf, pp, bw, ep["mpf_steps"]


def dual_step(w, mpf, pp, bw, mpf_steps, state):
    sv, model = w["svmpc"], w["model"]
    dyn = mpf.prior
    sv.optimize(state, dyn)
    a_seq, _ = sv.forward(state, dyn)
    action = a_seq[0]
    nxt = model.step(state.view(1, -1), action.view(1, -1), pp).view(-1)
    mpf.optimize(action.squeeze(), nxt.clone(), bw=bw, n_steps=mpf_steps)
    return nxt


def main():
    out = {"host": platform.processor() or platform.machine(), "cpus": os.cpu_count(), "torch": torch.__version__,
           "where": "build container (the reference cannot travel to the GPU box)", "configs": {}}
    for kind in ("pendulum", "particle"):
        best = None
        for th in (1, 2, 4, 8):
            if th > (os.cpu_count() or 1):
                continue
            torch.set_num_threads(th)
            w, mpf, pp, bw, ms = build(kind)
            state = w["state"].clone()
            for _ in range(2):
                dual_step(w, mpf, pp, bw, ms, state)
            n, t0 = 0, time.perf_counter()
            while n < 8:
                dual_step(w, mpf, pp, bw, ms, state)      # the state is held fixed: a timing loop, not an episode
                n += 1
            dt = (time.perf_counter() - t0) * 1e3 / n
            print(kind, "threads", th, f"{dt:.1f} ms per dual step", flush=True)
            if best is None or dt < best[1]:
                best = (th, dt)
        out["configs"][kind + "_demo"] = {"ms_per_dual_step": best[1], "dual_steps_per_sec": 1e3 / best[1], "threads": best[0],
                                          "kind": "reference (unmodified, via oracle/refshim.py stand-ins for gpytorch/KDEpy)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_reference_cpu_demo.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
