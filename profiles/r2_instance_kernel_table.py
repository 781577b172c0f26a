"""profiles/r2_instance_kernel_ncu.md from the ncu reports of the fused instance kernel (scratch copies under gpurun_out/).
usage: python profiles/r2_instance_kernel_table.py"""
import csv, subprocess

REPS = [("round 1: `svmpc_instance_kernel<pendulum,20,2>` (CTA tiles, accumulators in registers)", "gpurun_out/fused_r8.ncu-rep"),
        ("round 2a: `svmpc_warp_kernel` (per-warp TMA tiles, accumulators in shared memory, pair/item tail)", "gpurun_out/fused_r2a.ncu-rep"),
        ("round 2b: + host-side coefficients (no FP64), one buffer per warp, templated fold", "gpurun_out/fused_r2b.ncu-rep"),
        ("round 2c: + packed terminal cost, MUFU soft-min weights, running tile pointer (fewer instructions, same duration: "
         "the FMA-heavy pipe -- every packed FFMA2/FADD2/FMUL2 holds it two cycles -- and the issue slots bind)", "gpurun_out/fused_r2c.ncu-rep"),
        ("round 2e (off by default, DUST_B200_QUAD=1): `svmpc_quad_kernel`, two pairs per lane, 4 CTAs/SM -- fewer instructions, two "
         "independent chains per warp, and slower: 0.320 vs 0.309 ms in the same bench run (`r2_bench_run25_quad/pair.json`)", "gpurun_out/fused_r2e.ncu-rep")]  # fused_r2d.ncu-rep: the same kernel recaptured after common.cuh changed (traffic.json)
KEYS = [("gpu__time_duration.sum", "duration (us, under ncu)"), ("smsp__inst_executed.sum", "warp instructions"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (shared-memory limit)"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA (KB)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active (%)"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active (%)"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe cycles active (% of elapsed)"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active (%)"),
        ("dram__bytes_read.sum", "DRAM read (MB)"), ("dram__bytes_write.sum", "DRAM written (MB)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)")]
STEPS = 4096 * 2048 * 20  # trajectory-steps of the bench launch

cols = []
for title, rep in REPS:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    cols.append((title, d))
lines = ["# Round 2 - the fused instance kernel before / after (ncu --set full, bench launch: 4096 instances x 8 x 256 x H=20)", "",
         "| metric | " + " | ".join(f"({i + 1})" for i in range(len(cols))) + " |", "|---|" + "---|" * len(cols)]
for key, name in KEYS:
    lines.append(f"| {name} | " + " | ".join(str(d.get(key, "-")) for _, d in cols) + " |")
lines.append("| issue slots per trajectory-step | " + " | ".join(f"{float(d['smsp__inst_executed.sum']) * 32 / STEPS:.1f}" for _, d in cols) + " |")
st = [k for k in cols[0][1] if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
for k in st:
    vals = []
    for _, d in cols:
        try:
            vals.append(float(d.get(k, "0")))
        except ValueError:
            vals.append(0.0)
    if max(vals) > 0.3:
        lines.append(f"| stall `{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}` (warps per issue) | "
                     + " | ".join(f"{v:.2f}" for v in vals) + " |")
lines += [""] + [f"({i + 1}) {t}  [`{r}`]" for i, (t, r) in enumerate(REPS)]
open("profiles/r2_instance_kernel_ncu.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
