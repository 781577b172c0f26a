#!/usr/bin/env python
"""Large-N SVGD phi microbenchmark (BASELINE.json configs[3]): N particles, d = 40, canonical SVGD
direction with the exact median bandwidth.  Reports time, algorithmic TFLOP/s (6 N^2 d for phi,
4 N^2 d for the two median passes) and, under torchrun, the row-block sharded version
(all-gather of [X|score] + histogram all-reduce over NCCL).

  python bench_phi.py --particles 65536 --steps 5
  python -m torch.distributed.run --nproc-per-node 8 bench_phi.py --particles 65536
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=65536)
    ap.add_argument("--dim", type=int, default=40)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="compare a row sample against the float64 oracle")
    ap.add_argument("--gather", default="packed", choices=["packed", "separate"],
                    help="sharded runs: one all-gather of [X | score] (default) or X and score gathered separately")
    ap.add_argument("--emulate-world", type=int, default=1,
                    help="single process only: time the row block ONE rank of a world of this size computes (all N "
                         "columns resident, no collective) -- the per-rank device work of the sharded run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist

    from dust_b200 import _lib as L
    from dust_b200 import ops
    from dust_b200.distributed import ShardedSVGD, row_block

    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N, D = args.particles, args.dim
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(N, D, device=dev, generator=g)
    S = -X
    emu = args.emulate_world if world == 1 else 1
    b, e = row_block(N, rank, world) if emu == 1 else row_block(N, emu // 2, emu)
    sh = ShardedSVGD(N, D, device=dev, gather=args.gather)
    xl, sl = X[b:e].contiguous(), S[b:e].contiguous()
    if emu > 1:   # this process plays rank emu // 2: its row block against the resident columns
        sh.rows = (b, e)
        sh.gather = lambda x_local, s_local: (X, S)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    coef_holder = {}

    def full():
        phi, coef = sh.phi(xl, sl)
        coef_holder["coef"], coef_holder["phi"] = coef, phi

    if emu > 1:
        # one rank's histogram alone selects nothing (the all-reduce over the other ranks is missing):
        # take the bandwidth from a full single-GPU pass, untimed, and time only the phi row block
        med = ops.median_sq_dist(X)
        coef_full = ops.bandwidth_from_median(med, N, 1.0, 0)

        def full():  # noqa: F811
            x_all, s_all = sh.gather(xl, sl)
            out_ = ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma_dev=coef_full, rows=sh.rows)
            coef_holder["coef"], coef_holder["phi"] = coef_full, out_["phi"][0, b:e]

    ms_full = timed(full)
    coef = coef_holder["coef"]
    bw = float(coef[3])

    def phi_only():
        x_all, s_all = sh.gather(xl, sl)
        ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma_dev=coef, rows=sh.rows)

    ms_phi = timed(phi_only)
    lib = L.load()
    lib.dust_profiler_reset(); lib.dust_profiler_enable(1)
    full()
    prof = L.profiler_report()
    lib.dust_profiler_enable(0)
    # roofline denominator: measured cuBLAS TF32 GEMM throughput on this GPU (8192^3, best of 5)
    tf32_peak = None
    if rank == 0:
        torch.backends.cuda.matmul.allow_tf32 = True
        A_ = torch.randn(8192, 8192, device=dev); B_ = torch.randn(8192, 8192, device=dev)
        for _ in range(2):
            A_ @ B_
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); A_ @ B_; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf32_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del A_, B_
    out = None
    if rank == 0:
        fl_phi, fl_med = 6.0 * N * N * D / emu, 4.0 * N * N * D / emu
        Dp, NV = (D + 7) // 8 * 8, (2 * D + 15) // 16 * 16
        issued = 3 * 2.0 * N * N * (Dp + NV) / (world * emu)   # 3xTF32: hi*hi + hi*lo + lo*hi, per GPU
        t_phi_kernel = prof.get("phi_tc_kernel", (0, 0.0))[1]
        out = {"metric": "svgd_phi_large_n", "N": N, "d": D, "n_gpus": world, "emulated_world": emu, "rows": [b, e], "gather": args.gather,
               "ms_phi_with_median": ms_full if emu == 1 else None, "ms_phi": ms_phi,
               "algorithmic_tflops_phi": fl_phi / (ms_phi * 1e-3) / 1e12,
               "algorithmic_tflops_with_median": (fl_phi + fl_med) / (ms_full * 1e-3) / 1e12 if emu == 1 else None, "bandwidth": bw,
               "kernels_ms": {k: v[1] for k, v in prof.items()},
               "roofline": None if not t_phi_kernel else {
                   "kernel": "phi_tc_kernel", "bound": "tensor", "unit": "TFLOP/s",
                   "achieved": issued / (t_phi_kernel * 1e-3) / 1e12, "peak": tf32_peak,
                   "frac": issued / (t_phi_kernel * 1e-3) / 1e12 / tf32_peak,
                   "peak_source": "cuBLAS TF32 GEMM 8192^3 measured in this run (torch.matmul, allow_tf32)",
                   "note": "achieved counts the TF32 MMA FLOPs actually issued (3 per algorithmic product)"}}
    if args.check:
        from oracle import dust_oracle as O
        idx = torch.arange(b, e, max(1, (e - b) // 64), device=dev)[:64]
        Xc, Sc = X.cpu().double(), S.cpu().double()
        g_, c1, c2 = [float(v) for v in coef[:3].cpu()]
        xi = Xc[idx.cpu()]
        d2 = ((xi * xi).sum(-1, keepdim=True) + (Xc * Xc).sum(-1)[None, :] - 2 * xi @ Xc.t()).clamp(min=0)
        K = (-g_ * d2).exp()
        ref = c1 * (K @ Sc) + c2 * (K.sum(1, keepdim=True) * xi - K @ Xc)
        got = coef_holder["phi"][idx - b].cpu().double()
        err = float((got - ref).abs().max() / ref.abs().max())
        if rank == 0:
            out["rel_err_vs_float64_rows"] = err
            if N <= 16384:
                bw_ref, _ = O.bw_median(X.cpu())
                out["bandwidth_ref"] = float(bw_ref)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
