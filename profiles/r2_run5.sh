# round 2, fifth hardware run (one GPU): cluster kernel for few instances, one-launch optimize in the classes, the parameter
# filter without host round trips (device Silverman, ParticleBelief, 16 lanes per particle); demo-shape timings
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "not 65536" --durations=6 > gpurun_out/pytest_gpu_r2_run5.log 2>&1; tail -n 30 gpurun_out/pytest_gpu_r2_run5.log
timeout -s KILL 300 python bench_configs.py --steps 30 --warmup 5 > gpurun_out/bench_configs_r2_run5.json 2> gpurun_out/bench_configs_r2_run5.err
python - <<'PY'
import json
try:
    for l in open("gpurun_out/bench_configs_r2_run5.json"):
        if not l.startswith("{"): continue
        d = json.loads(l)
        if "config" not in d: print(d); continue
        print(d["config"], "device", round(d["device_ms_per_dual_step"], 4), "wall", round(d["wall_ms_per_dual_step"], 4), "launches", d["library_launches_per_step"], "classes", d.get("drop_in_classes"))
        print("   ", {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print("failed", e)
print(open("gpurun_out/bench_configs_r2_run5.err").read()[-1500:])
PY
timeout -s KILL 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -q -x -k "cluster_kernel or silverman or device_bandwidth" > gpurun_out/sanitize_cluster_memcheck.log 2>&1; tail -n 5 gpurun_out/sanitize_cluster_memcheck.log
timeout -s KILL 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_round2.py -q -x -k "cluster_kernel and pendulum-1-3" > gpurun_out/sanitize_cluster_racecheck.log 2>&1; tail -n 5 gpurun_out/sanitize_cluster_racecheck.log
