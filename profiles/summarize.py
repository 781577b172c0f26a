"""Turn the scratch captures under gpurun_out/ into the tracked summaries under profiles/.
usage: python profiles/summarize.py   (after `gpurun -- 'bash profiles/r1_collect.sh'`)"""
import collections, csv, json, re, shutil, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2]

def table(rep, want, stalls=True):
    hdr, units, vals = raw(rep)
    lines = ["| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hdr:
            i = hdr.index(w); lines.append(f"| `{w}` | {vals[i]} | {units[i]} |")
    if stalls:
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                if v > 0.1:
                    lines.append(f"| `{h}` | {v:.3f} | warps per issue |")
    get = lambda w: float(vals[hdr.index(w)].replace(",", "")) if w in hdr else float("nan")
    return lines, get

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def launch_list(src, dst_md, dst_csv, title, notes):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [title, "", "`ncu --metrics gpu__time_duration.sum --clock-control none -c 200` (per-launch times are cold-cache and serialised: read the SHARES).",
           f"Raw csv: `{dst_csv}`.  Collected by `profiles/r1_collect.sh`, summarised by `profiles/summarize.py`.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n[:140]}` | {a[0]} | {a[1]:.1f} | {a[1]/a[0]:.1f} | {100*a[1]/tot:.1f}% |")
    out += [""] + notes
    open(dst_md, "w").write("\n".join(out) + "\n")
    shutil.copy(src, dst_csv)
    return agg

if __name__ == "__main__":
    agg = launch_list("gpurun_out/launches_r1b.csv", "profiles/r1_launch_list.md", "profiles/r1_launches.csv",
                      "# Round 1 - ncu launch list of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline` (B200, batched pendulum SVMPC)",
                      ["Reading: the capture covers the whole process (value pass, profiler pass, e2e pass).  In the value pass a control step is ONE launch of",
                       "`svmpc_instance_kernel` (rollouts + costs + soft-min + likelihood gradient + GMM prior score + RBF phi + SGD update + weights / argmax /",
                       "shift / prior refresh); `gpu_launches` = steps.  `noise_normal_kernel` is the library's Philox/Box-Muller generator drawing the 671 MB noise",
                       "tensor: outside the timed `value` region (noise resident), inside the `e2e` region (one draw per call).  The `at::` kernels are torch's",
                       "set-up of the synthetic problem and the state/action staging copies.  `bench.py`'s live CUDA-event profile gives the step kernel",
                       "99 % of the `value` step, as here (no other `dust::` kernel runs in that region)."])
    lines, get = table("gpurun_out/fused_r8.ncu-rep", WANT)
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    md = ["# Round 1 - `ncu --set full` of the dominant kernel `svmpc_instance_kernel<pendulum, ACC=20, TPT=2>` (one-launch control step)", "",
          "Workload: bench.py default (4096 instances x 8 policies x 256 samples x H=20, P=1), one launch, `--clock-control none`.",
          "The launch is the WHOLE control step: TMA bulk fetch of the noise tiles, two trajectories per thread on the packed FP32 pipe",
          "(FFMA2/FADD2/FMUL2), costs + online soft-min + analytic likelihood gradient, then the tail (GMM prior score, RBF phi among the",
          "8 policies, SGD update, weights / argmax / shift / prior refresh).",
          "Report: `gpurun_out/fused_r8.ncu-rep` (scratch; command in `profiles/r1_collect.sh`).", ""] + lines + ["",
          f"DRAM traffic per launch = {rd:.1f} MB read + {wr:.1f} MB written = {rd+wr:.1f} MB; algorithmic bytes = 707.3 MB (noise 671.1 + costs 33.6 +",
          "theta/state 2.7) + ~8 MB of tail outputs that bench.py's formula leaves out: no re-reads.", "",
          "History of this kernel in round 1 (same workload, ncu `smsp__inst_executed.sum` / duration):",
          "", "| state | warp instructions | duration |", "|---|---|---|",
          "| rollouts only, 4 launches per step (`fused_r4`) | 481 M | 489 us (+149 us in three more kernels) |",
          "| one-launch step, scalar (`fused_r5`) | 527 M | 536 us |",
          "| shorter trig reduction, contracted dynamics, sigma*eps rows (`fused_r6`) | 461 M | ~475 us |",
          "| + TMA bulk tile fetch, one exp per thread in the combine | - | ~415 us |",
          "| + two trajectories per thread on FFMA2 (`fused_r7`), 4 CTAs/SM | 292 M | 398 us |",
          f"| + 5 CTAs/SM, single-copy fallback (`fused_r8`, this table) | {get('smsp__inst_executed.sum')/1e6:.0f} M | {get('gpu__time_duration.sum'):.0f} us |", "",
          "The packed path removed 37 % of the issued instructions but issue efficiency fell from 89 % to ~2/3: with ~100 registers per thread",
          "only 20 warps fit an SM and the dependent chain of a model step (reduction -> polynomial -> select -> dynamics) is no longer hidden.",
          "Bound now: latency at 20 warps/SM (FMA pipe ~55 %, ALU ~46 %); HBM stays at ~1/4 of the copy bandwidth."]
    open("profiles/r1_fused_instance_kernel_ncu.md", "w").write("\n".join(md) + "\n")
    json.dump({"kernel": "svmpc_instance_kernel<pendulum,20,2> (one-launch control step)",
               "source": "profiles/r1_fused_instance_kernel_ncu.md (ncu --set full, one launch, bench workload, capture fused_r8)",
               "dram_bytes_read": rd * 1e6, "dram_bytes_write": wr * 1e6, "dram_bytes_per_launch": (rd + wr) * 1e6},
              open("profiles/rollout_traffic.json", "w"), indent=1)
    lines, get = table("gpurun_out/noise_r1.ncu-rep", WANT)
    n_bytes = 4096 * 256 * 8 * 20 * 4 / 1e6
    md = ["# Round 1 - `ncu --set full` of `noise_normal_kernel` (Philox4x32-10 + Box-Muller action noise)", "",
          f"One fill of the bench's noise tensor: {n_bytes:.1f} MB written once.  Report: `gpurun_out/noise_r1.ncu-rep`.", ""] + lines + ["",
          f"Algorithmic bytes = {n_bytes:.1f} MB (write only) -> {n_bytes/get('gpu__time_duration.sum')*1e3/1e3:.2f} TB/s at this duration."]
    open("profiles/r1_noise_kernel_ncu.md", "w").write("\n".join(md) + "\n")
    for f, t in (("bench_r1_b.json", "r1_bench.json"), ("bench_r1_ref.json", "r1_bench_ref.json")):
        shutil.copy("gpurun_out/" + f, "profiles/" + t)
    print(open("profiles/r1_launch_list.md").read()[:1800])
