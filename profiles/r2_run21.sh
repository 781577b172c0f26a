# round 2, twenty-first hardware run (one GPU): the whole GPU suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl
timeout -s KILL 1500 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/pytest_gpu_r2_run21.log 2>&1; tail -n 14 gpurun_out/pytest_gpu_r2_run21.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2_run21.log 2>&1; tail -n 3 gpurun_out/smoke_r2_run21.log
( time timeout -s KILL 600 python bench.py > gpurun_out/bench_r2_run21.json 2> gpurun_out/bench_r2_run21.err ) 2>&1 | grep real
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2_run21.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e ms", d["e2e"]["ms_per_step"], d["clocks"])
    p = d["phi"]; print("phi", p["ms_phi"], p["ms_phi_with_median"], p["roofline"]["frac"], p["roofline"]["frac_vs_inrun_cublas_tf32"], p["roofline"].get("traffic"), p["rel_err_vs_float64_rows"], p["median"]["ulp_distance"], p["clocks"])
    print("phi kernels", {k: round(v, 4) for k, v in p["kernels_ms"].items()})
    for n, c in d["configs"].items():
        if n == "clocks": continue
        print(n, "device", round(c["device_ms_per_dual_step"], 4), "wall", round(c["wall_ms_per_dual_step"], 4), "launches", c["library_launches_per_step"], (c.get("drop_in_classes") or {}).get("wall_ms_per_dual_step"))
    print("cpu", d["cpu_baseline"]["value"], {k: round(v["ms_per_dual_step"], 2) for k, v in d["cpu_baseline"].get("configs", {}).items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_run21.err").read()[-2500:])
PY
( time timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_ref_r2_run21.json 2> gpurun_out/bench_ref_r2_run21.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_ref_r2_run21.json
