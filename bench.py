#!/usr/bin/env python
"""bench.py -- DuSt-MPC control steps/s on B200 (BASELINE.json metric).

Workload (configs[2], the configuration the 1/2/4/8-GPU metric is quoted on): batched pendulum
SVMPC, 4096 independent MPC instances PER GPU x 8 policies x 256 action samples x horizon 20,
no parameter sampling (P = 1).  One "step" = SVMPC.optimize (n_steps = 1) + SVMPC.forward for
every instance: GMM prior score -> rollout + cost + soft-min likelihood gradient -> fused
SVGD phi + SGD update -> weights / argmax / shift / prior refresh.

  python bench.py --gpus N --steps K --warmup W            (one rank per GPU under torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle port of the
                                                             reference's eager-torch path)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(N=8, S=256, H=20, A=1, ds=2, P=1, ctrl_sigma=2.0, prior_sigma=2.0, alpha=1.0, lr=2.0)
METRIC, UNIT = "svmpc_control_steps_per_sec", "instance-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=4096, help="MPC instances per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="instances in the bounded CPU sample (default: sized for ~15 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-phi", action="store_true", help="skip the large-N SVGD phi block (configs[3])")
    ap.add_argument("--no-configs", action="store_true", help="skip the demo / dual-stress block (configs[0], [1], [4])")
    ap.add_argument("--phi-steps", type=int, default=20)
    ap.add_argument("--config-steps", type=int, default=40)
    return ap.parse_args()


def workload_config(args, n_gpus):
    c = CFG
    return {
        "workload": "batched pendulum SVMPC (BASELINE.json configs[2])",
        "instances_per_gpu": args.instances, "policies": c["N"], "action_samples": c["S"], "horizon": c["H"],
        "params_samples": c["P"], "kernel": "rbf (gpytorch default lengthscale)",
        "likelihood": "ExponentiatedUtility", "sharding": f"instances x{n_gpus}, no data-path collective",
        "rollouts_per_step": args.instances * n_gpus * c["N"] * c["S"] * c["P"],
        "l2_policy": "noise input (671 MB/GPU/step) is larger than the 126 MB L2",
    }


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path, one instance at a time (the reference has no
# batch dimension: B instances are B sequential SVMPC.optimize + SVMPC.forward calls)
# ---------------------------------------------------------------------------------------------
def cpu_instance_steps(n_instances, n_steps, threads, seed=0):
    from oracle import dust_oracle as O

    c = CFG
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    model = O.Model("pendulum")
    sigma = torch.tensor([c["ctrl_sigma"]])
    insts = []
    for _ in range(n_instances):
        mu = torch.randn(c["N"], c["H"], c["A"], generator=g)
        theta = mu + c["prior_sigma"] * torch.randn(c["N"], c["H"], c["A"], generator=g)
        st = O.SvmpcState(theta, mu, torch.ones(c["N"]), c["prior_sigma"] ** 2)
        state = (torch.rand(2, generator=g) * 2 - 1) * torch.tensor([3.14159, 1.0])
        insts.append((st, state))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        for st, state in insts:
            eps = torch.randn(c["S"], c["N"], c["H"], c["A"], generator=g)  # reference draws by rsample
            out = O.svmpc_optimize(model, st, state, eps, sigma, None, False, c["alpha"], c["lr"], kernel="rbf")
            O.svmpc_forward(st, out["costs"], c["alpha"], False)
    dt = time.perf_counter() - t0
    return n_instances * n_steps / dt, dt


def cpu_demo_steps(threads, seconds=6.0):
    """The two demo configurations (configs[0], [1]) on the CPU port: one dual control step = SVMPC optimize +
    forward, plant step, 20 MPF steps (particle_example.py:177-207).  -> {name: ms per dual step}."""
    from bench_configs import CONFIGS, PARTICLE_ENV
    from dust_b200.models.particle import Particle
    from oracle import dust_oracle as O

    torch.set_num_threads(threads)
    out = {}
    for name in ("pendulum_demo", "particle_demo"):
        c = CONFIGS[name]
        g = torch.Generator().manual_seed(0)
        if c["kind"] == "pendulum":
            model, A, state = O.Model("pendulum"), 1, torch.tensor([3.0, 0.0])
            params = 0.6 + 0.7 * torch.rand(c["P"], 2, generator=g)
            x = 0.6 + 0.7 * torch.rand(c["Np"], 2, generator=g)
        else:
            model, A, state = O.Model("particle", O.ParticleCfg(Particle(**PARTICLE_ENV, uncertain_params=["mass"], mass=2.0).obst_map.map)), 2, torch.tensor([-9.0, -9.0, 0.0, 0.0])
            params = 0.693 + 0.1 * torch.randn(c["P"], 1, generator=g)
            x = 0.693 + 0.1 * torch.randn(c["Np"], 1, generator=g)
        N, H, S = c["N"], c["H"], c["S"]
        mu = torch.randn(N, H, A, generator=g)
        st = O.SvmpcState(mu + c["prior_sigma"] * torch.randn(N, H, A, generator=g), mu, torch.ones(N), c["prior_sigma"] ** 2)
        sigma = torch.full((A,), c["sigma"])

        def dual():
            nonlocal x
            eps = torch.randn(S, N, H, A, generator=g)
            o = O.svmpc_optimize(model, st, state, eps, sigma, params, c["log_space"], 1.0, c["lr"], kernel="rbf")
            a_seq, _, _ = O.svmpc_forward(st, o["costs"], 1.0, c["weighted_prior"])
            act = a_seq[0]
            nxt = model.step(state.reshape(1, -1), act.reshape(1, -1)).reshape(-1)
            x, _ = O.mpf_optimize(model, x, state, act, nxt, c["obs_std"], torch.full((model.dp,), c["mpf_bw"] ** 2), c["mpf_bw"],
                                  c["mpf_lr"], c["mpf_steps"], c["log_space"])

        dual()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds and n < 200:
            dual()
            n += 1
        ms = (time.perf_counter() - t0) * 1e3 / n
        out[name] = {"ms_per_dual_step": ms, "dual_steps_per_sec": 1e3 / ms, "steps_timed": n, "cores": threads, "kind": "port"}
    return out


def probe_threads():
    """best intra-op thread count for the eager-torch port on this host (tiny tensors: more
    threads are often slower)"""
    ncpu = os.cpu_count() or 1
    probe = {}
    for th in sorted({1, min(2, ncpu), min(4, ncpu), min(8, ncpu), min(16, ncpu)}):
        cpu_instance_steps(2, 1, th)  # warm-up
        probe[th] = cpu_instance_steps(24, 1, th)[0]
    return max(probe, key=probe.get), probe


def best_cpu(n_instances, target_seconds=15.0):
    th, probe = probe_threads()
    if n_instances <= 0:
        n_instances = min(4096, max(64, int(probe[th] * target_seconds)))
    v, dt = cpu_instance_steps(n_instances, 1, th)
    return v, th, dt, n_instances, probe


def run_reference(args, rank):
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    # thread sweep on a small probe, then K timed "steps", each a bounded sample of n instances
    th, probe = probe_threads()
    n = args.cpu_sample if args.cpu_sample > 0 else min(args.instances * args.gpus, max(16, int(probe[th] * 60.0 / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        cpu_instance_steps(max(2, n // 8), 1, th)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_instance_steps(n, 1, th, seed=k)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} of {args.instances * args.gpus} instances per step, one at a time (the reference has no batch dim)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port", "sample": sample,
                         "host_cpus": ncpu, "thread_probe": {str(k): v for k, v in probe.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
from bench_common import ClockSampler, committed_traffic, measured_peaks  # noqa: E402

ROLLOUT_SOURCES = ("dust_b200/csrc/rollout.cu", "dust_b200/csrc/models.cuh", "dust_b200/csrc/common.cuh")


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from dust_b200 import _lib as L
    from dust_b200.batched import BatchedSVMPC
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    L.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    c = CFG
    B = args.instances
    ctl = BatchedSVMPC(PendulumModel(), B, c["N"], c["S"], c["H"], c["ctrl_sigma"], c["prior_sigma"], alpha=c["alpha"],
                       learning_rate=c["lr"], kernel="gpytorch", inst_cost_fn=inst_cost, term_cost_fn=term_cost,
                       device=dev, seed=1234 + rank, prefetch_noise=not os.environ.get("DUST_BENCH_NO_PREFETCH"))
    g = ctl.gen
    state = (torch.rand(B, 2, device=dev, generator=g) * 2 - 1) * torch.tensor([3.14159, 1.0], device=dev)
    eps = ctl.draw_noise()  # resident in HBM for the `value` measurement
    lib = L.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        ctl.control_step(state, eps)   # SVMPC.optimize + SVMPC.forward for every instance

    for _ in range(args.warmup):
        step_resident()
    # ---- value: K steps, inputs resident in HBM ------------------------------------------------
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.15)
    n0 = lib.dust_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = int(lib.dust_launch_count() - n0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel breakdown with the library's event timer (same steps, second pass) -------
    lib.dust_profiler_reset()
    lib.dust_profiler_enable(1)
    for _ in range(args.steps):
        step_resident()
    prof = L.profiler_report()
    lib.dust_profiler_enable(0)

    # ---- e2e: host state in, action out, noise drawn on the device by the public API ----------
    h_state = state.cpu().pin_memory()
    h_action = torch.empty(B, c["A"]).pin_memory()
    d_state = torch.empty_like(state)
    for _ in range(max(3, args.warmup // 2)):
        d_state.copy_(h_state, non_blocking=True)
        h_action.copy_(ctl.control_step(d_state), non_blocking=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        d_state.copy_(h_state, non_blocking=True)
        h_action.copy_(ctl.control_step(d_state), non_blocking=True)
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms_total = float(ms2[0])

    # ---- the other BASELINE configurations, in the same process (every rank takes part in the collectives) ----
    del eps, ctl
    torch.cuda.empty_cache()
    phi = cfgs = None
    if not args.no_phi:
        from bench_phi import phi_block

        phi = phi_block(rank, world, dev, steps=args.phi_steps, warmup=3)
        torch.cuda.empty_cache()
    if not args.no_configs:
        from bench_configs import configs_block

        cfgs = configs_block(rank, world, dev, steps=args.config_steps, warmup=3)
    if rank != 0:
        return
    K = args.steps
    total_inst = B * world
    value = total_inst * K / (ms_total * 1e-3)
    e2e_value = total_inst * K / (e2e_ms_total * 1e-3)
    # roofline of the dominant kernel (rollout + cost): algorithmic bytes per launch =
    # 4 * B * (S*N*H*A noise read + S*N costs written + N*H*A theta read + ds state read)
    algo_bytes = 4.0 * B * (c["S"] * c["N"] * c["H"] * c["A"] + c["S"] * c["N"] + c["N"] * c["H"] * c["A"] + c["ds"])
    peaks = measured_peaks()
    peak, peak_src = peaks["hbm_gbs"], peaks["source"] + " hbm_gbs (measured copy bandwidth)"
    dom = "svmpc_instance_kernel" if "svmpc_instance_kernel" in prof else "rollout_cost_kernel"  # fused step kernel
    n_roll, ms_roll = prof.get(dom, (0, 0.0))
    roof = None
    if n_roll:
        per = ms_roll / n_roll
        ach = algo_bytes / (per * 1e-3) / 1e9
        # ncu dram__bytes_read + write of this kernel, committed with the hash of the sources it was captured on
        traffic, traffic_note = committed_traffic(dom, ROLLOUT_SOURCES)
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_note": traffic_note, "algorithmic_bytes_per_launch": algo_bytes,
                "ms_per_launch": per, "peak_source": peak_src,
                "note": "issue-bound, not HBM-limited (~30 issue slots per model step against 4 bytes of noise): see DESIGN.md section 3"}
    step_kernel_ms = {k: v[1] / K for k, v in prof.items()}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, th, dt, n_cpu, probe = best_cpu(args.cpu_sample)
        cpu = {"value": v, "unit": UNIT, "cores": th, "kind": "port", "host_cpus": os.cpu_count(),
               "thread_probe": {str(k): round(x, 1) for k, x in probe.items()},
               "sample": f"{n_cpu} of {B} instances, 1 step each, one instance at a time ({dt:.1f} s of CPU work)"}
        if cfgs is not None:
            cpu["configs"] = cpu_demo_steps(th)
    cfg = workload_config(args, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "rollouts_per_sec": value * c["N"] * c["S"] * c["P"],
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_total / K,
                "h2d_bytes_per_step": B * c["ds"] * 4, "d2h_bytes_per_step": B * c["A"] * 4,
                "note": "state in (pinned host) -> action out (pinned host); noise drawn on the device inside the call"
                        + ("" if os.environ.get("DUST_BENCH_NO_PREFETCH") else "; the draw of step k + 1 is enqueued on a side stream while step k runs (BatchedSVMPC(prefetch_noise=True))")},
        "gpu_launches": launches, "kernel_ms_per_step": step_kernel_ms, "roofline": roof, "cpu_baseline": cpu,
        "clocks": clocks, "phi": phi, "configs": cfgs,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
