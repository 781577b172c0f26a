"""profiles/r2_phi_a_in_tmem.md: phi_tc_kernel before / after the round-2 changes, from the ncu reports (scratch copies under
gpurun_out/; N = 65536, d = 40, both launches of one call).  usage: python profiles/r2_phi_kernel_table.py"""
import csv
import subprocess

REPS = [("round 2c: single issuing warp, row tile (A of GEMM1) in shared memory (SS form), row-major unit ranges for the rest", "gpurun_out/phi_r2c.ncu-rep"),
        ("round 2d: A staged in TMEM by the flush warps (TS-form GEMM1), the rest in column chunks", "gpurun_out/phi_r2d.ncu-rep"),
        ("round 2e: + two issuing warps (GEMM1 / GEMM2), S_EMPTY barrier", "gpurun_out/phi_r2e.ncu-rep"),
        ("round 2f: + GEMM2's P_lo V term as kind::f16 on bf16 copies (P_lo: 32 TMEM columns per buffer), S/P_hi ring of three", "gpurun_out/phi_r2f.ncu-rep"),
        ("round 2g (withdrawn): softmax loop on the packed FP32 pipe -- fewer instructions, 128 registers, latency-bound, slower", "gpurun_out/phi_r2g.ncu-rep"),
        ("round 2h: scalar softmax loop with an explicit fma, kernel templated on its form (no runtime mode branches, no predicated-off duplicates)", "gpurun_out/phi_r2h.ncu-rep"),
        ("round 2i: + the flush warps sleep between polls of O_FULL (final)", "gpurun_out/phi_r2i.ncu-rep")]
KEYS = [("gpu__time_duration.sum", "duration (ms, under ncu)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "`sm__pipe_tensor_cycles_active` (% of elapsed)"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "the `TriageCompute ... realtime` variant of it (%; sampled, unstable)"),
        ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA (KB)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active (%)"),
        ("dram__bytes_read.sum", "DRAM read (MB)"), ("dram__bytes_write.sum", "DRAM written (MB)"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU shared-memory wavefronts (% of peak)")]
UNITS = {0: 444 * 1024, 1: 68 * 1024}   # tile pairs of the two launches


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return [dict(zip(rows[0], r)) for r in rows[2:]]


cols = [(t, load(r), r) for t, r in REPS]
lines = ["# Round 2 - `phi_tc_kernel` before / after (ncu --set full --clock-control none, N = 65536, d = 40)", "",
         "Two launches per call: 444 whole row tiles (3 per SM, all CTAs sweep the columns together), then the other 68 row tiles.",
         "Cells: first launch / second launch.", "",
         "| metric | " + " | ".join(f"({i + 1})" for i in range(len(cols))) + " |", "|---|" + "---|" * len(cols)]
for key, name in KEYS:
    lines.append(f"| {name} | " + " | ".join(" / ".join(str(d.get(key, "-")) for d in ds) for _, ds, _ in cols) + " |")
lines.append("| SM cycles per tile pair (1965 MHz; tensor work 15 x 32 + 24 x 40 = 1440 cycles, 1280 from (4) on: 4 bf16 MMAs replace 8 TF32) | " + " | ".join(
    " / ".join(f"{float(d['gpu__time_duration.sum']) * 1e-3 * 1.965e9 / (UNITS[i] / 148):.0f}" for i, d in enumerate(ds)) for _, ds, _ in cols) + " |")
lines.append("| tensor pipe active, both launches weighted by duration (%) | " + " | ".join(
    f"{sum(float(d['gpu__time_duration.sum']) * float(d[KEYS[1][0]]) for d in ds) / sum(float(d['gpu__time_duration.sum']) for d in ds):.1f}"
    for _, ds, _ in cols) + " |")
lines += [""] + [f"({i + 1}) {t}  [`{r}`]" for i, (t, _, r) in enumerate(cols)]
lines += ["",
          "Reading (per-instruction samples of the same reports, `--page source`): in (1) the issuing warp never waited for P "
          "(0.3 % of its samples on `P_FULL`) while the softmax warps spent 35 % of theirs waiting for S: the tensor pipe was starved "
          "by its feeder, not by the exponentials.  Two causes.  (a) An SS-form 128x64x8 TF32 MMA reads 4 KB of A and 2 KB of B from "
          "shared memory in the 32 cycles it occupies the pipe - 192 B/cycle against the 128 B/cycle the shared-memory port delivers - "
          "so GEMM1 ran at 48 cycles per MMA; with the row tile in TMEM (80 of the 96 free columns) only B is fetched.  (b) One warp "
          "issued all 39 MMAs of a tile pair plus ~250 descriptor moves, waits and commits, ~2500 cycles per tile pair against 1440 "
          "cycles of tensor work; two issuing warps halve that.  In (3) the GEMM1 warp waits for `S_EMPTY` 47 % of its time and the "
          "softmax warps wait for S 23 % of theirs: the chain S -> exp -> P -> GEMM2 of the two S buffers is what remains "
          "(a third buffer no longer fits beside A: 192 + 128 + 160 + 80 > 512 columns).",
          "(4) makes room for it: the P_lo V term is 2^-11 of the sum, so P_lo = bf16(K - K_hi) and a bf16 copy of V go through "
          "`kind::f16` (16 elements of K per instruction, two values per TMEM column: 192 + 64 + 160 + 80 = 496 columns).  The bf16 pair "
          "order in a TMEM column (even element in the low half) was settled by measurement: the other order is 1.7e-4 from float64, "
          "this one 4e-6 like the all-TF32 form (`profiles/r2_run18.sh`).  After (4) the softmax warps hardly wait (S_FULL 3 % of their "
          "samples) and are issue-bound (418 instructions per tile and warp, `selected` + `not_selected` 47 %).  (5) cut that to 323 with "
          "FFMA2 / FADD2 / FMUL2 and LOST: 128 registers, the |x_j|^2 loads no longer hoisted (`short_scoreboard` 23 %, `wait` 34 %), the "
          "GEMM2 warp back to waiting for P 37 % of its time.  (6) keeps the scalar loop, writes the fma explicitly (ptxas had emitted "
          "s + s, two adds), and templates the kernel on its form so that neither the TF32 split nor the other form's MMAs are issued "
          "predicated-off: 96 registers.",
          "Tried after (7) and withdrawn: the |x_j|^2 slices in a ring of their own (eight stages, filled by the X producer, recycled by the "
          "softmax warps) so that the softmax warps do not wait for the V stage that carried them: 2.74 -> 2.85 ms on the main launch "
          "(`gpurun_out/phi_r2k.ncu-rep`, scratch) -- the single producer thread now paces the X tiles.",
          "Two counters: `sm__pipe_tensor_cycles_active` (the one BASELINE names) is stable from capture to capture (86.1 / 86.4 % for the "
          "identical kernels of (6) and (7)); the `TriageCompute ... realtime` variant that round 1 quoted (54 %) is sampled and is not "
          "(80.4 vs 53.4 % for the same two captures): it is listed for continuity only.  Neither is a pure work counter -- (4) does the "
          "same job in fewer tensor cycles AND less time -- so read them together with the duration row.",
          "Timed in `bench_phi.py` (steady state, power-capped clocks): 5.09 -> 4.70 -> 4.17 -> 3.85 -> (4.10) -> 3.53 ms of kernel per "
          "call, phi 5.05 -> 3.33 ms (`profiles/r2_bench_phi_run8.json`, `r2_bench_phi_run9_a_tmem_w1.json`, `r2_bench_phi_run10_w1.json`, "
          "`r2_bench_phi_run18_w1.json`, `r2_bench_phi_run19_w1.json`, `r2_bench_phi_run20_w1.json`); one rank's row block of 8: "
          "0.675 -> 0.437 ms."]
open("profiles/r2_phi_a_in_tmem.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
