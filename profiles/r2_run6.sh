# round 2, sixth hardware run (one GPU): whole GPU suite (the two 140-s N = 65536 cases excepted: run in r2_run1), demo shapes, bench line
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -k "not 65536" --durations=5 > gpurun_out/pytest_gpu_r2_run6.log 2>&1; tail -n 25 gpurun_out/pytest_gpu_r2_run6.log
timeout -s KILL 500 python bench.py > gpurun_out/bench_r2_run6.json 2> gpurun_out/bench_r2_run6.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2_run6.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e ms", d["e2e"]["ms_per_step"])
    p = d["phi"]; print("phi", p["ms_phi"], p["ms_phi_with_median"], p["roofline"]["frac"], p["roofline"]["frac_vs_inrun_cublas_tf32"], p["rel_err_vs_float64_rows"], p["median"]["ulp_distance"], p["clocks"])
    print("phi kernels", {k: round(v, 4) for k, v in p["kernels_ms"].items()})
    for n, c in d["configs"].items():
        if n == "clocks": continue
        print(n, "device", round(c["device_ms_per_dual_step"], 4), "wall", round(c["wall_ms_per_dual_step"], 4), "launches", c["library_launches_per_step"], c.get("drop_in_classes"))
        print("   ", {k: round(v, 4) for k, v in c["kernel_ms_per_step"].items()})
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("configs"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_run6.err").read()[-2500:])
PY
timeout -s KILL 200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_round2.py -q -x -k "cluster_kernel and pendulum-1-3" > gpurun_out/sanitize_cluster_racecheck.log 2>&1; tail -n 3 gpurun_out/sanitize_cluster_racecheck.log
