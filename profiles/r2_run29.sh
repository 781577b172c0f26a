# round 2, twenty-ninth hardware run (one GPU): the whole GPU suite on the final sources, smoke, bench line, captures for traffic.json
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl
timeout -s KILL 1500 python -m pytest tests -q -m gpu --durations=4 > gpurun_out/pytest_gpu_r2_run29.log 2>&1; tail -n 9 gpurun_out/pytest_gpu_r2_run29.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2_run29.log 2>&1; tail -n 2 gpurun_out/smoke_r2_run29.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_r2_run29.json 2> gpurun_out/bench_r2_run29.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2_run29.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e ms", d["e2e"]["ms_per_step"], d["clocks"])
    p = d["phi"]; print("phi", p["ms_phi"], p["ms_phi_with_median"], p["roofline"]["frac"], p["roofline"]["frac_vs_inrun_cublas_tf32"], p["roofline"].get("traffic"), p["rel_err_vs_float64_rows"], p["median"]["ulp_distance"], p["clocks"])
    print("phi kernels", {k: round(v, 4) for k, v in p["kernels_ms"].items()})
    for n, c in d["configs"].items():
        if n == "clocks": continue
        print(n, "device", round(c["device_ms_per_dual_step"], 4), "graph", (c.get("cuda_graph") or {}).get("device_ms_per_dual_step"), "classes", (c.get("drop_in_classes") or {}).get("wall_ms_per_dual_step"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_run29.err").read()[-2500:])
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 4 -c 2 -o gpurun_out/phi_r2j -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_phi_r2j.log 2>&1; tail -n 1 gpurun_out/ncu_phi_r2j.log
