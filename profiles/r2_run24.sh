# round 2, twenty-fourth hardware run (one GPU): the demo dual step as a CUDA graph
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_round2.py -q -x -k "captured" 2>&1 | tail -n 3
timeout -s KILL 300 python bench_configs.py --configs pendulum_demo,particle_demo --steps 40 > gpurun_out/bench_configs_r2_run24.json 2> gpurun_out/bench_configs_r2_run24.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs_r2_run24.json"):
    if not l.startswith("{"): continue
    d = json.loads(l)
    if "config" in d: print(d["config"], "device", round(d["device_ms_per_dual_step"], 4), "wall", round(d["wall_ms_per_dual_step"], 4), "graph", d.get("cuda_graph"), "classes", (d.get("drop_in_classes") or {}).get("wall_ms_per_dual_step"))
PY
tail -n 5 gpurun_out/bench_configs_r2_run24.err
