# round 2, twenty-sixth hardware run (one GPU): new tests (quad kernel, captured step), bench line, instance-kernel capture on the final sources
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_round2.py -q -x -k "two_pairs or captured or peer or packed_buffer" 2>&1 | tail -n 3
timeout -s KILL 600 python bench.py > gpurun_out/bench_r2_run26.json 2> gpurun_out/bench_r2_run26.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2_run26.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e ms", d["e2e"]["ms_per_step"], d["clocks"])
    p = d["phi"]; print("phi", p["ms_phi"], p["ms_phi_with_median"], p["roofline"]["frac"], p["roofline"].get("traffic"), p["rel_err_vs_float64_rows"], p["median"]["ulp_distance"])
    for n, c in d["configs"].items():
        if n == "clocks": continue
        print(n, "device", round(c["device_ms_per_dual_step"], 4), "graph", (c.get("cuda_graph") or {}).get("device_ms_per_dual_step"), "classes", (c.get("drop_in_classes") or {}).get("wall_ms_per_dual_step"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_run26.err").read()[-2500:])
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:svmpc_warp_kernel -s 3 -c 1 -o gpurun_out/fused_r2f -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-phi --no-configs > gpurun_out/ncu_fused_r2f.log 2>&1; tail -n 1 gpurun_out/ncu_fused_r2f.log
