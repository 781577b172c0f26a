"""Second round-1 collection -> tracked summaries.  usage: python profiles/summarize2.py
(after `gpurun --timeout 900 -- 'bash profiles/r1_collect2.sh'`; reads gpurun_out/, writes profiles/)."""
import csv, json, shutil, subprocess

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")].split("(")[0]
        lines = ["| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in hdr:
                lines.append(f"| `{w}` | {vals[hdr.index(w)]} | {units[hdr.index(w)]} |")
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                if v > 0.1:
                    lines.append(f"| stall `{h.split('stalled_')[1].split('_per')[0]}` per issue | {v:.3f} | |")
        yield name, lines


def first_json(path):
    return [json.loads(l) for l in open(path) if l.startswith("{")]


if __name__ == "__main__":
    md = ["# Round 1, second collection - `ncu --set full` of the tensor-core kernels (N = 65536, d = 40, one launch each, `--clock-control none`)", "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:\"phi_tc_kernel|median_tc_kernel\" -c 2 python bench_phi.py --steps 1 --warmup 0`",
          "(`profiles/r1_collect2.sh`; report `gpurun_out/tc_r4.ncu-rep`, scratch).  `phi_tc_kernel` is the persistent version: 148 CTAs, each an equal",
          "contiguous range of (row tile, column tile) pairs.  SASS evidence (cuobjdump): `UTCHMMA` (tcgen05.mma kind::tf32, SS and TS forms),",
          "`LDTM`/`STTM` (tcgen05.ld/st), `UBLKCP` (cp.async.bulk), `UTCBAR` (tcgen05.commit).", ""]
    for name, lines in kernels("gpurun_out/tc_r4.ncu-rep"):
        md += [f"## `{name}`", ""] + lines + [""]
    open("profiles/r1_tensor_core_kernels_ncu.md", "w").write("\n".join(md))
    md = ["# Round 1, second collection - `ncu --set full` of `rollout_cost_kernel<particle, plain, NSUB = 4>` at the dual-stress shape", "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:rollout_cost_kernel -c 1 python bench_configs.py --configs dual_stress --steps 1 --warmup 0`.",
          "P = 512 draws x S = 1024 x N = 32 x H = 50 = 839 M model steps per launch; four groups of 128 threads share one 52 KB action tile", ""]
    for name, lines in kernels("gpurun_out/rollout_r1.ncu-rep"):
        md += [f"## `{name}`", ""] + lines + [""]
    open("profiles/r1_rollout_cost_kernel_ncu.md", "w").write("\n".join(md))
    shutil.copy("gpurun_out/bench_r1_c.json", "profiles/r1_bench.json")
    shutil.copy("gpurun_out/bench_r1_ref_c.json", "profiles/r1_bench_ref.json")
    shutil.copy("gpurun_out/bench_phi_r1_c.json", "profiles/r1_bench_phi.json")
    with open("profiles/r1_bench_phi_row_blocks.json", "w") as f:   # one rank's row block of a world of 2 / 4 / 8, timed on one GPU
        for w in (2, 4, 8):
            for d in first_json(f"gpurun_out/bench_phi_r1_c_emu{w}.json"):
                f.write(json.dumps(d) + "\n")
    with open("profiles/r1_bench_configs.json", "w") as f:
        for name in ("bench_configs_r1_c.json", "bench_configs_r1_c_dense.json", "bench_configs_r1_c_emu8.json"):
            for d in first_json("gpurun_out/" + name):
                f.write(json.dumps(d) + "\n")
    shutil.copy("gpurun_out/pytest_gpu_r1c.log", "profiles/r1_pytest_gpu.log")
    print(open("profiles/r1_tensor_core_kernels_ncu.md").read()[:3000])
    print(open("profiles/r1_rollout_cost_kernel_ncu.md").read()[:3000])
