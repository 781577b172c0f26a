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
    import os

    import torch.distributed as dist

    from dust_b200 import _lib as L
    from dust_b200.distributed import ShardedRollout, row_block

    L.require_cuda()
    rank, world, lrank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1:   # one instance, its parameter draws split over the ranks (inputs replicated: same seed everywhere)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for name in args.configs.split(","):
        cfg = CONFIGS[name]
        p = build(cfg, dev, alpha=args.alpha)
        if world > 1 or args.emulate_world > 1:
            sh = ShardedRollout(cfg["P"])
            if args.emulate_world > 1:
                sh.p_range = row_block(cfg["P"], args.emulate_world // 2, args.emulate_world)
            p["core"].sharded = sh
        for _ in range(args.warmup):
            dual_step(cfg, p)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            dual_step(cfg, p)
        e1.record()
        sync()
        wall_ms = (time.perf_counter() - w0) * 1e3 / args.steps
        dev_ms = e0.elapsed_time(e1) / args.steps
        if world > 1:   # device time of the slowest rank
            t = torch.tensor([dev_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms = float(t[0])
        lib = L.load()
        lib.dust_profiler_reset()
        lib.dust_profiler_enable(1)
        for _ in range(args.steps):
            dual_step(cfg, p)
        torch.cuda.synchronize()
        prof = {k: v[1] / args.steps for k, v in L.profiler_report().items()}
        lib.dust_profiler_enable(0)
        model_steps = cfg["P"] * cfg["S"] * cfg["N"] * cfg["H"]
        line = {"config": name, "shape": {k: cfg[k] for k in ("kind", "H", "N", "S", "P", "Np", "mpf_steps", "grad")},
                "device_ms_per_dual_step": dev_ms, "wall_ms_per_dual_step": wall_ms,
                "dual_steps_per_sec": 1e3 / dev_ms, "rollouts_per_sec": cfg["P"] * cfg["S"] * cfg["N"] * 1e3 / dev_ms,
                "model_steps_per_control_step": model_steps, "kernel_ms_per_step": prof, "steps": args.steps,
                "warmup": args.warmup, "data": "synthetic", "alpha": args.alpha, "n_gpus": world,
                "emulated_world": args.emulate_world,
                "sharding": None if p["core"].sharded is None else "parameter draws %s of %d on this rank" % (
                    list(p["core"].sharded.p_range), cfg["P"])}
        lw = p["core"].last.get("lik_weights")
        if lw is not None:   # pathwise gradient: rows with an exactly-zero weight are not rolled out by the adjoint
            line["nonzero_weight_fraction"] = float((lw != 0).float().mean())
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
