#!/usr/bin/env python
"""Per-configuration timing of ONE dual control step (SVMPC optimize + forward, then MPF) on one
B200, for the BASELINE.json configurations that `bench.py` does not quote its headline on:

  pendulum_demo  configs[0]  H=30 N=3  S=128  P=8   MPF Np=50  x 20 steps  (pendulum_config.yaml)
  particle_demo  configs[1]  H=40 N=6  S=64   P=4   MPF Np=50  x 20 steps  (particle_config.yaml)
  dual_stress    configs[4]  H=50 N=32 S=1024 P=512 MPF Np=512 x 20 steps, pathwise (adjoint) gradient

These are single-instance problems (B = 1): the demo shapes are launch/latency bound, the stress
shape is the one that exercises the adjoint kernel at scale.
One JSON line per configuration: device ms per step (CUDA events over the timed steps), wall ms per
step (host overhead included), per-kernel ms from the library profiler and rollouts per second.
The reference's own CPU timings of the two demo shapes are in SURVEY.md section 6 / BASELINE.md.

`configs_block()` is what bench.py embeds in its JSON line: the kernel-level step (`SvmpcCore` + ops, device and wall
ms) AND, for the two demo shapes, the same dual step through the drop-in classes a user of the reference calls
(`SVMPC.optimize` / `SVMPC.forward` / `MPF.optimize`, wall clock, launches per step).  Under torchrun the stress shape's
parameter draws are split over the ranks (`ShardedRollout`, NCCL all-reduce of the cost and gradient shares) and rank
0 also runs the unsharded step in the same process: strong scaling and a parity figure.

Synthetic inputs, seeded; noise resident on the device.  Not a parity test (tests/ hold those)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARTICLE_ENV = dict(dt=0.015, control_type="acceleration", noise_std=[0.1, 0.1], init_state=[-9.0, -9.0, 0, 0],
                    target_state=[9.0, 9.0, 0, 0], can_crash=True, with_obstacle=True, deterministic=True,
                    cost_params=dict(w_qpos=0.5, w_qvel=0.25, w_ctrl=0.2, w_obs=1.0e6, w_qpos_T=1.0e3, w_qvel_T=0.1),
                    obst_preset="grid_4x4", obst_width=2.1, max_speed=5, max_accel=10, map_cell_size=0.1,
                    map_size=[22, 22], map_type="direct")

CONFIGS = {
    "pendulum_demo": dict(kind="pendulum", H=30, N=3, S=128, P=8, Np=50, mpf_steps=20, sigma=2.0, prior_sigma=2.0, lr=2.0,
                          grad="analytic", log_space=False, mpf_lr=1e-3, mpf_bw=0.05, obs_std=0.1, weighted_prior=False),
    "particle_demo": dict(kind="particle", H=40, N=6, S=64, P=4, Np=50, mpf_steps=20, sigma=5.0, prior_sigma=5.0, lr=100.0,
                          grad="analytic", log_space=True, mpf_lr=1e-2, mpf_bw=0.5, obs_std=0.1, weighted_prior=True),
    "dual_stress": dict(kind="particle", H=50, N=32, S=1024, P=512, Np=512, mpf_steps=20, sigma=5.0, prior_sigma=5.0, lr=100.0,
                        grad="pathwise", log_space=True, mpf_lr=1e-2, mpf_bw=0.5, obs_std=0.1, weighted_prior=True),
}


def build(cfg, dev, seed=0, alpha=1.0):
    from dust_b200.inference.core import SvmpcCore
    from dust_b200.models.particle import Particle
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    if cfg["kind"] == "pendulum":
        model = PendulumModel(uncertain_params=("length", "mass"))
        spec = model.device_spec(inst_cost, term_cost, dev)
        state = torch.tensor([[3.0, 0.0]], device=dev)
        params = 0.6 + 0.7 * torch.rand(1, cfg["P"], 2, device=dev, generator=g)          # U([.6,.6],[1.3,1.3])
        mpf_x = 0.6 + 0.7 * torch.rand(1, cfg["Np"], 2, device=dev, generator=g)
        A = 1
    else:
        model = Particle(**PARTICLE_ENV, uncertain_params=["mass"], mass=2.0)
        spec = model.device_spec(model.default_inst_cost, model.default_term_cost, dev)
        state = torch.tensor([[-9.0, -9.0, 0.0, 0.0]], device=dev)
        params = torch.exp(torch.log(torch.tensor(2.0)) + 0.1 * rn(1, cfg["P"], 1))      # masses (already exp'd)
        mpf_x = torch.log(torch.tensor(2.0)) + 0.1 * rn(1, cfg["Np"], 1)                 # log-space particles
        A = 2
    N, H, S = cfg["N"], cfg["H"], cfg["S"]
    mu = rn(1, N, H, A)
    theta = mu + cfg["prior_sigma"] * rn(1, N, H, A)
    core = SvmpcCore(spec, theta, mu, torch.ones(1, N, device=dev), torch.full((A,), cfg["prior_sigma"] ** 2),
                     torch.full((A,), cfg["sigma"]), alpha=alpha, temperature=1.0 / alpha, lr=cfg["lr"], kernel="gpytorch",
                     grad=cfg["grad"], weighted_prior=cfg["weighted_prior"])
    eps = rn(1, S, N, H, A)
    return dict(spec=spec, core=core, state=state.contiguous(), params=params.contiguous(), eps=eps, mpf_x=mpf_x.contiguous(), A=A)


def dual_step(cfg, p):
    """optimize + forward, plant step with the chosen action, MPF on the observed transition."""
    from dust_b200 import ops

    core, spec = p["core"], p["spec"]
    core.optimize_step(p["state"], p["eps"], p["params"])
    a_seq, _, _ = core.forward_step()
    action = a_seq[:, 0].contiguous()
    nxt = ops.model_step(spec, p["state"], action)
    inv_var = torch.full((spec.dp,), 1.0 / cfg["mpf_bw"] ** 2, device=nxt.device)
    ops.mpf_optimize(spec, p["mpf_x"], p["state"], action, nxt, inv_var, cfg["obs_std"], cfg["mpf_bw"], cfg["mpf_lr"],
                     cfg["mpf_steps"], cfg["log_space"])
    return nxt


def dual_step_inplace(cfg, p):
    """`dual_step` with the controller's state held in PERSISTENT buffers (what a CUDA graph needs: `SvmpcCore` rebinds
    theta / mix to fresh tensors every step, a captured graph would keep reading the old ones): the particles and the
    mixture weights are pointed at the buffers before the step and copied back into them after it.  Steady state only
    (the prior already aliases the particles)."""
    core = p["core"]
    if "theta_p" not in p:
        p["theta_p"], p["mix_p"] = core.theta.clone(), core.mix.clone()
    core.theta = core.mu = p["theta_p"]
    core.mix = p["mix_p"]
    nxt = dual_step(cfg, p)
    p["theta_p"].copy_(core.theta)
    p["mix_p"].copy_(core.mix)
    core.theta = core.mu = p["theta_p"]
    core.mix = p["mix_p"]
    return nxt


def graph_dual_step(cfg, p, steps, warmup):
    """The same dual step captured ONCE into a CUDA graph and replayed (the C ABI allocates nothing and only enqueues on
    the stream it is given, so its launches are capturable; torch's few helper kernels are captured with them).
    -> {"device_ms_per_dual_step", "wall_ms_per_dual_step", "nodes"} or {"unavailable": why}."""
    try:
        from dust_b200.utils.graphs import CapturedStep

        g = CapturedStep(lambda: dual_step_inplace(cfg, p)).graph
        for _ in range(warmup):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return {"device_ms_per_dual_step": e0.elapsed_time(e1) / steps, "wall_ms_per_dual_step": (time.perf_counter() - w0) * 1e3 / steps,
                "note": "one torch.cuda.CUDAGraph holding the control step, the plant step, the 20 MPF steps and the two copies that "
                        "keep the controller state in persistent buffers; replayed (state evolves across replays: "
                        "tests/test_gpu_round2.py::test_captured_control_step_replays_the_eager_sequence)"}
    except Exception as exc:  # noqa: BLE001
        torch.cuda.synchronize()
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def class_objects(name, dev, seed=0):
    """The reference-shaped objects of a demo configuration (demo/*_example.py), on the device."""
    import torch.distributions as dist

    from demo import configs
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import ExponentiatedUtility, GaussianLikelihood
    from dust_b200.inference.mpf import MPF
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBFKernel
    from dust_b200.models.pendulum import PendulumModel
    from dust_b200.models.pendulum import inst_cost as pend_inst
    from dust_b200.models.pendulum import term_cost as pend_term

    torch.manual_seed(seed)
    if name == "particle_demo":
        from demo.particle_example import build as build_particle

        cfg = configs.load(None, configs.PARTICLE)
        controller, svmpc, mpf, model, _ = build_particle(cfg, seed)
        ep = cfg["exp_params"]
        state = torch.as_tensor(cfg["env_params"]["init_state"], dtype=torch.float).to(dev)
        return dict(svmpc=svmpc, mpf=mpf, plant=model, state=state, mpf_bw=ep["mpf_bandwidth"], mpf_steps=ep["mpf_steps"])
    cfg = configs.load(None, configs.PENDULUM)
    ep = cfg["exp_params"]
    H, N, S, A = ep["horizon"], ep["n_particles"], ep["action_samples"], ep["ctrl_dim"]
    model = PendulumModel(uncertain_params=("length", "mass"))
    prior = get_gmm(torch.randn(N, H, A), torch.ones(N), ep["prior_sigma"] ** 2 * torch.eye(A))
    theta0 = prior.sample([N])
    dyn_prior = dist.Independent(dist.Uniform(torch.tensor([0.6, 0.6]), torch.tensor([1.3, 1.3])), 1)
    ctrl = MultiDISCO(observation_space=model.observation_space, action_space=model.action_space, hz_len=H, n_policies=N,
                      action_samples=S, params_samples=ep["params_samples"], temperature=1 / ep["alpha"],
                      a_cov=ep["ctrl_sigma"] ** 2 * torch.eye(A), inst_cost_fn=pend_inst, term_cost_fn=pend_term,
                      params_sampling=True, params_log_space=ep["mpf_log_space"])
    lik = ExponentiatedUtility(ep["alpha"], n_samples=S, controller=ctrl, model=model)
    svmpc = SVMPC(init_particles=theta0, prior=prior, likelihood=lik, kernel=RBFKernel(), n_particles=N,
                  bw_scale=ep["bandwidth_scaling"], n_steps=1, optimizer_class=torch.optim.SGD, lr=ep["learning_rate"])
    state = torch.as_tensor(ep["init_state"], dtype=torch.float).to(dev)
    mpf_init = dyn_prior.sample([ep["mpf_n_particles"]])
    dyn_lik = GaussianLikelihood(initial_obs=state, obs_std=ep["mpf_obs_std"], model=model, log_space=ep["mpf_log_space"])
    mpf = MPF(init_particles=mpf_init, likelihood=dyn_lik, optimizer_class=torch.optim.SGD, lr=ep["mpf_learning_rate"],
              bw=ep["mpf_bandwidth"], bw_scale=ep["mpf_bandwidth_scaling"])
    plant = PendulumModel(length=1.0, mass=1.0)
    return dict(svmpc=svmpc, mpf=mpf, plant=plant, state=state, mpf_bw=ep["mpf_bandwidth"], mpf_steps=ep["mpf_steps"])


def class_dual_step(o):
    """One control step exactly as the reference's drivers make it (particle_example.py:177-207,
    simulations.py:104-138): optimise, act, step the plant, condition the parameter filter."""
    sv, mpf = o["svmpc"], o["mpf"]
    dyn = mpf.prior
    sv.optimize(o["state"], dyn)
    a_seq, _ = sv.forward(o["state"], dyn)
    action = a_seq[0]
    nxt = o["plant"].step(o["state"].reshape(1, -1), action.reshape(1, -1)).reshape(-1)
    mpf.optimize(action.squeeze(), nxt, bw=o["mpf_bw"], n_steps=o["mpf_steps"])
    return nxt


def time_class_path(name, dev, steps, warmup):
    from dust_b200 import _lib as L

    lib = L.load()
    o = class_objects(name, dev)
    for _ in range(warmup):
        class_dual_step(o)       # the state is held fixed: a timing loop, not an episode
    torch.cuda.synchronize()
    n0 = lib.dust_launch_count()
    w0 = time.perf_counter()
    for _ in range(steps):
        class_dual_step(o)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1e3 / steps
    ref = None
    try:   # the UNMODIFIED reference timed on the build container's cores (profiles/reference_cpu_demo.py; it cannot travel)
        r = json.load(open(os.path.join(ROOT, "profiles", "r2_reference_cpu_demo.json")))
        ref = {"ms_per_dual_step": r["configs"][name]["ms_per_dual_step"], "threads": r["configs"][name]["threads"],
               "host_cpus": r["cpus"], "where": r["where"], "ratio": r["configs"][name]["ms_per_dual_step"] / wall_ms}
    except Exception:
        pass
    return {"wall_ms_per_dual_step": wall_ms, "dual_steps_per_sec": 1e3 / wall_ms, "reference_unmodified_cpu": ref,
            "library_launches_per_step": (lib.dust_launch_count() - n0) / steps,
            "api": "SVMPC.optimize + SVMPC.forward + model.step + MPF.optimize (drop-in classes, belief = mpf.prior)"}


def configs_block(rank, world, dev, names=("pendulum_demo", "particle_demo", "dual_stress"), steps=10, warmup=3, alpha=1.0,
                  emulate_world=1, class_path=True):
    """-> {config name: timing dict} on rank 0 (None elsewhere)."""
    import torch.distributed as dist

    from bench_common import ClockSampler
    from dust_b200 import _lib as L
    from dust_b200.distributed import ShardedRollout, row_block

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {}
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    for name in names:
        cfg = CONFIGS[name]
        if world > 1 and name != "dual_stress":
            continue           # the demo shapes are single-instance, single-GPU problems ("replicas only")
        p = build(cfg, dev, alpha=alpha)
        if world > 1 or emulate_world > 1:
            sh = ShardedRollout(cfg["P"])
            if emulate_world > 1:
                sh.p_range = row_block(cfg["P"], emulate_world // 2, emulate_world)
            p["core"].sharded = sh
        for _ in range(warmup):
            dual_step(cfg, p)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            dual_step(cfg, p)
        e1.record()
        sync()
        wall_ms = (time.perf_counter() - w0) * 1e3 / steps
        dev_ms = e0.elapsed_time(e1) / steps
        if world > 1:   # device time of the slowest rank
            t = torch.tensor([dev_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms = float(t[0])
        lib = L.load()
        lib.dust_profiler_reset()
        lib.dust_profiler_enable(1)
        n0 = lib.dust_launch_count()
        for _ in range(steps):
            dual_step(cfg, p)
        torch.cuda.synchronize()
        launches = (lib.dust_launch_count() - n0) / steps
        prof = {k: v[1] / steps for k, v in L.profiler_report().items()}
        lib.dust_profiler_enable(0)
        model_steps = cfg["P"] * cfg["S"] * cfg["N"] * cfg["H"]
        line = {"shape": {k: cfg[k] for k in ("kind", "H", "N", "S", "P", "Np", "mpf_steps", "grad")},
                "device_ms_per_dual_step": dev_ms, "wall_ms_per_dual_step": wall_ms,
                "dual_steps_per_sec": 1e3 / dev_ms, "rollouts_per_sec": cfg["P"] * cfg["S"] * cfg["N"] * 1e3 / dev_ms,
                "model_steps_per_control_step": model_steps, "library_launches_per_step": launches,
                "kernel_ms_per_step": prof, "steps": steps, "warmup": warmup, "alpha": alpha, "n_gpus": world,
                "emulated_world": emulate_world,
                "sharding": None if p["core"].sharded is None else "parameter draws %s of %d on rank 0" % (
                    list(p["core"].sharded.p_range), cfg["P"])}
        lw = p["core"].last.get("lik_weights")
        if lw is not None:   # pathwise gradient: rows with an exactly-zero weight are not rolled out by the adjoint
            line["nonzero_weight_fraction"] = float((lw != 0).float().mean())
        if world > 1:
            # the same step unsharded on rank 0 alone (same process, same inputs): strong scaling + parity of one step
            q = build(cfg, dev, alpha=alpha)
            single_ms = None
            if rank == 0:
                for _ in range(warmup):
                    dual_step(cfg, q)
                torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(steps):
                    dual_step(cfg, q)
                f1.record()
                torch.cuda.synchronize()
                single_ms = f0.elapsed_time(f1) / steps
            sync()
            a, b = build(cfg, dev, alpha=alpha), build(cfg, dev, alpha=alpha)
            a["core"].sharded = ShardedRollout(cfg["P"])
            a["core"].optimize_step(a["state"], a["eps"], a["params"])       # collective: every rank
            if rank == 0:
                b["core"].optimize_step(b["state"], b["eps"], b["params"])
                rel = lambda x, y: float((x - y).abs().max() / y.abs().max())  # noqa: E731
                line["single_gpu_same_process_ms"] = single_ms
                line["strong_scaling"] = single_ms / (world * dev_ms)
                line["max_rel_diff_vs_single_gpu"] = {"costs": rel(a["core"].last["costs"], b["core"].last["costs"]),
                                                      "grad_lik": rel(a["core"].last["grad_lik"], b["core"].last["grad_lik"]),
                                                      "theta": rel(a["core"].theta, b["core"].theta)}
                line["collectives"] = {"costs_all_reduce_bytes": 4 * cfg["S"] * cfg["N"],
                                       "gradient_all_reduce_bytes": 4 * cfg["N"] * cfg["H"] * p["A"] if cfg["grad"] == "pathwise" else 0}
            sync()
        if class_path and world == 1 and emulate_world == 1 and name != "dual_stress":
            line["drop_in_classes"] = time_class_path(name, dev, steps, warmup)
            line["cuda_graph"] = graph_dual_step(cfg, p, steps, warmup)      # last: a failed capture must not disturb the rest
        out[name] = line
    if sampler:
        out["clocks"] = sampler.stop()
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="pendulum_demo,particle_demo,dual_stress")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--alpha", type=float, default=1.0, help="likelihood temperature (demos: 1); a tiny value makes every soft-min weight non-zero")
    ap.add_argument("--emulate-world", type=int, default=1,
                    help="single process: time the parameter-draw share ONE rank of a world of this size rolls out "
                         "(no all-reduce: the device work of a rank, not a valid control step)")
    args = ap.parse_args()
    import torch.distributed as dist

    from dust_b200 import _lib as L

    L.require_cuda()
    rank, world, lrank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1:   # one instance, its parameter draws split over the ranks (inputs replicated: same seed everywhere)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    out = configs_block(rank, world, dev, tuple(args.configs.split(",")), args.steps, args.warmup, args.alpha, args.emulate_world)
    if rank == 0:
        for name, line in out.items():
            print(json.dumps({"config": name, **line} if isinstance(line, dict) and name != "clocks" else {name: line}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
