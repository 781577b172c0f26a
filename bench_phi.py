#!/usr/bin/env python
"""Large-N SVGD phi (BASELINE.json configs[3]): N particles, d = 40, canonical SVGD direction
(dust/inference/svgd.py:127-135) with the exact median bandwidth (svgd.py:42-52).

`phi_block()` is what bench.py embeds in its JSON line on every run: time, tensor-pipe roofline of
`phi_tc_kernel` (issued 3xTF32 FLOPs against bf16_tflops/2 of MEASURED_PEAKS.json AND against cuBLAS TF32
measured in the same run), its own clock sample, and in-run checks (sampled rows against float64, the
tensor-core median against the SIMT radix select).  Under torchrun the rows are sharded over the ranks
(`ShardedSVGD`: all-gather of X and score, all-reduce of the bandwidth histogram, NCCL), then rank 0 times the
full problem alone in the same process: `strong_scaling = t1 / (n * tn)` and `max_rel_diff_vs_single_gpu`.

  python bench_phi.py --particles 65536 --steps 5
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_phi.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench_common import ClockSampler, committed_traffic, cublas_tf32_tflops, measured_peaks, timed_ms  # noqa: E402

PHI_SOURCES = ["dust_b200/csrc/svgd_tc.cu", "dust_b200/csrc/common.cuh"]


def float64_rows(X, S, idx, gamma, c1, c2):
    """phi rows `idx` in float64 with torch on the device: c1 K S + c2 (rowsum(K) x - K X), K = exp(-gamma d^2)
    (the unified form of svgd.py:127-135)."""
    Xd, Sd = X.double(), S.double()
    xi = Xd[idx]
    d2 = ((xi * xi).sum(-1, keepdim=True) + (Xd * Xd).sum(-1)[None, :] - 2 * xi @ Xd.t()).clamp(min=0)
    d2[torch.arange(len(idx), device=X.device), idx] = 0.0
    K = (-gamma * d2).exp()
    return c1 * (K @ Sd) + c2 * (K.sum(1, keepdim=True) * xi - K @ Xd)


def phi_block(rank, world, dev, steps=5, warmup=3, N=65536, D=40, gather="peer", emulate_world=1, checks=True):
    import torch.distributed as dist

    from dust_b200 import _lib as L
    from dust_b200 import ops
    from dust_b200.distributed import ShardedSVGD, row_block

    lib = L.load()
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(N, D, device=dev, generator=g)       # same seed on every rank: each holds the full cloud,
    S = -X                                               # uses only its row block as its local input
    emu = emulate_world if world == 1 else 1
    b, e = row_block(N, rank, world) if emu == 1 else row_block(N, emu // 2, emu)
    gather_note = None
    if world > 1 and gather == "peer":
        # the NVLink peer-memory exchange needs CUDA IPC between the ranks' processes: if the box refuses it, every rank
        # learns so here (the decision is all-reduced) and the job takes the NCCL all-gather instead -- and says so
        ok = torch.ones(1, device=dev)
        try:
            from dust_b200.distributed import PeerExchange
            probe = PeerExchange(N // world, 2 * D, device=dev)
        except Exception as exc:  # noqa: BLE001
            ok.zero_()
            gather_note = f"peer exchange unavailable ({type(exc).__name__}: {exc}); NCCL all-gather used"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            gather = "packed"
            gather_note = gather_note or "peer exchange unavailable on another rank; NCCL all-gather used"
    sh = ShardedSVGD(N, D, device=dev, gather=gather)
    if world > 1 and gather == "peer":
        sh._peer = probe
    xl, sl = X[b:e].contiguous(), S[b:e].contiguous()
    if emu > 1:   # this process plays rank emu // 2: its row block against the resident columns
        sh.rows = (b, e)
        sh.gather = lambda x_local, s_local: (X, S)
    hold = {}

    def sharded_full():
        hold["phi"], hold["coef"] = sh.phi(xl, sl)

    if emu > 1:
        # one rank's histogram alone selects nothing (the all-reduce over the other ranks is missing): take the
        # bandwidth from a full single-GPU pass, untimed, and time only the phi row block
        coef_full = ops.bandwidth_from_median(ops.median_sq_dist(X), N, 1.0, 0)

        def sharded_full():  # noqa: F811
            x_all, s_all = sh.gather(xl, sl)
            out_ = ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma_dev=coef_full, rows=sh.rows)
            hold["phi"], hold["coef"] = out_["phi"][0, b:e], coef_full

    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    ms_full = timed_ms(sharded_full, steps, warmup, dev, world)
    coef = hold["coef"]

    def sharded_phi_only():
        x_all, s_all = sh.gather(xl, sl)
        ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma_dev=coef, rows=sh.rows)

    ms_phi = timed_ms(sharded_phi_only, steps, warmup, dev, world)
    lib.dust_profiler_reset(); lib.dust_profiler_enable(1)
    sharded_full()
    prof = L.profiler_report()
    lib.dust_profiler_enable(0)
    phi_sharded = hold["phi"].clone()

    single = None
    if world > 1:
        # rank 0 alone, full problem, same process and clocks; the other ranks wait at the barrier inside timed_ms
        def single_full():
            med = ops.median_sq_dist(X)
            c = ops.bandwidth_from_median(med, N, 1.0, 0)
            hold["phi1"], hold["coef1"] = ops.svgd_phi(X.unsqueeze(0), S.unsqueeze(0), gamma_dev=c)["phi"][0], c

        def single_phi_only():
            ops.svgd_phi(X.unsqueeze(0), S.unsqueeze(0), gamma_dev=coef)

        def noop():
            pass

        t1_full = timed_ms(single_full if rank == 0 else noop, steps, warmup, dev, world)
        t1_phi = timed_ms(single_phi_only if rank == 0 else noop, steps, warmup, dev, world)
        # parity of the sharded rows against the single-GPU rows (gathered on every rank; rank 0 compares)
        allphi = torch.empty(N, D, device=dev)
        dist.all_gather_into_tensor(allphi, phi_sharded.contiguous())
        single = {"t1_full": t1_full, "t1_phi": t1_phi}
        if rank == 0:
            ref = hold["phi1"]
            single["max_rel_diff"] = float((allphi - ref).abs().max() / ref.abs().max())
            single["bandwidth_bits_equal"] = bool(torch.equal(hold["coef1"], coef))
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return None

    peaks = measured_peaks()
    tf32_inrun = cublas_tf32_tflops(dev)
    Dp, NV = (D + 7) // 8 * 8, (2 * D + 15) // 16 * 16
    mode = int(lib.dust_phi_tc_mode(D))
    bf16lo = mode >= 0 and bool(mode & 2)
    # tensor work issued per GPU in TF32-equivalent FLOPs: GEMM1 3 TF32 MMAs per product (hi*hi + hi*lo + lo*hi); GEMM2
    # P_hi V_hi + P_hi V_lo in TF32 and the P_lo V correction either in TF32 or -- bf16lo -- as kind::f16 at twice the rate
    issued = 2.0 * N * N * (3 * Dp + (2.5 if bf16lo else 3.0) * NV) / (world * emu)
    n_phi, t_phi_kernel = prof.get("phi_tc_kernel", (0, 0.0))
    n_med, t_med_kernel = prof.get("median_tc_kernel", (0, 0.0))
    peak = peaks["bf16_tflops"] / 2.0
    fl_phi, fl_med = 6.0 * N * N * D, 4.0 * N * N * D
    out = {"workload": "large-N SVGD phi (BASELINE.json configs[3])", "N": N, "d": D, "n_gpus": world, "emulated_world": emu,
           "rows_of_rank0": [b, e], "gather": gather if world > 1 else None, "gather_note": gather_note, "steps": steps, "warmup": warmup,
           "ms_phi": ms_phi, "ms_phi_with_median": ms_full if emu == 1 else None,
           "algorithmic_tflops_phi": fl_phi / emu / (ms_phi * 1e-3) / 1e12,    # whole job (all ranks) over its time
           "algorithmic_tflops_with_median": (fl_phi + fl_med) / (ms_full * 1e-3) / 1e12 if emu == 1 else None,
           "bandwidth": float(coef[3]), "kernels_ms": {k: v[1] for k, v in prof.items()}, "clocks": clocks,
           "l2_policy": "operands (31 MB) are L2-resident by design: tensor-pipe bound, HBM traffic negligible"}
    if n_phi:
        ach = issued / (t_phi_kernel * 1e-3) / 1e12
        traffic, traffic_note = committed_traffic("phi_tc_kernel", PHI_SOURCES) if world * emu == 1 else (None, "captured for the single-GPU call only")
        out["roofline"] = {"kernel": "phi_tc_kernel", "bound": "tensor", "unit": "TFLOP/s", "achieved": ach, "peak": peak,
                           "frac": ach / peak, "peak_source": f"{peaks['source']} bf16_tflops / 2 (nominal TF32:BF16 = 1:2)",
                           "peak_inrun_cublas_tf32": tf32_inrun, "frac_vs_inrun_cublas_tf32": ach / tf32_inrun,
                           "ms_per_launch_sum": t_phi_kernel, "launches": n_phi, "traffic": traffic, "traffic_note": traffic_note,
                           "algorithmic_operand_bytes": 4.0 * N * (2 * Dp + 2 * NV + 1),
                           "kernel_form": {"row_tile_in_tmem": bool(mode & 1), "p_lo_term_bf16": bf16lo} if mode >= 0 else None,
                           "note": "achieved counts the tensor work issued in TF32-equivalent FLOPs (3 MMAs per Gram product, 2 TF32 + 1 "
                                   "bf16-at-half-cost or 3 TF32 per K V product; K padded to 8 / NV to 16); algorithmic FLOPs are "
                                   "6 N^2 d, about a third of it"}
    if n_med:
        issued_med = 3 * 2.0 * N * N * Dp / 2 / (world * emu)   # symmetric half band of tile pairs
        out["median_kernel"] = {"kernel": "median_tc_kernel", "ms": t_med_kernel,
                                "issued_tflops": issued_med / (t_med_kernel * 1e-3) / 1e12,
                                "frac_of_bf16_half": issued_med / (t_med_kernel * 1e-3) / 1e12 / peak}
    if single is not None:
        out["single_gpu_same_process"] = {"ms_phi": single["t1_phi"], "ms_phi_with_median": single["t1_full"]}
        out["strong_scaling"] = single["t1_phi"] / (world * ms_phi)
        out["strong_scaling_with_median"] = single["t1_full"] / (world * ms_full)
        out["max_rel_diff_vs_single_gpu"] = single["max_rel_diff"]
        out["bandwidth_bits_equal_to_single_gpu"] = single["bandwidth_bits_equal"]
    if checks and emu == 1:
        gam, c1, c2 = [float(v) for v in coef[:3].cpu()]
        gi = torch.Generator(device=dev).manual_seed(1)
        idx = torch.cat([torch.arange(b, e, max(1, (e - b) // 64), device=dev)[:64],
                         torch.randint(b, e, (64,), device=dev, generator=gi)])
        ref = float64_rows(X, S, idx, gam, c1, c2)
        got = phi_sharded[idx - b].double()
        out["rel_err_vs_float64_rows"] = float((got - ref).abs().max() / ref.abs().max())
        out["rows_checked"] = int(idx.numel())
        if world == 1:
            fast = ops.median_sq_dist(X).clone()
            robust = ops.median_sq_dist(X, allow_fast=False).clone()
            bits = lambda t: int(t.view(torch.int32)[0])  # noqa: E731
            out["median"] = {"tensor_core_window": float(fast[0]), "simt_radix_select": float(robust[0]),
                             "ulp_distance": abs(bits(fast) - bits(robust)),
                             "note": "two exact rank selections over differently rounded distances (3xTF32 Gram vs fp32 FMA chain)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=65536)
    ap.add_argument("--dim", type=int, default=40)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gather", default="peer", choices=["peer", "packed", "separate"],
                    help="sharded runs: rows pulled over NVLink peer memory (default; csrc/peer.cu), one NCCL all-gather of "
                         "[X | score] read in place through a row stride, or X and score gathered by two collectives")
    ap.add_argument("--emulate-world", type=int, default=1,
                    help="single process only: time the row block ONE rank of a world of this size computes (all N "
                         "columns resident, no collective) -- the per-rank device work of the sharded run")
    ap.add_argument("--no-checks", action="store_true", help="skip the float64 / radix-select checks (profiler captures)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist

    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    out = phi_block(rank, world, dev, args.steps, args.warmup, args.particles, args.dim, args.gather, args.emulate_world,
                    checks=not args.no_checks)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
