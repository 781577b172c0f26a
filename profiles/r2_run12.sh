# round 2, twelfth hardware run (one GPU): instance kernel per-tile trims (packed terminal cost, MUFU weights, running tile pointer)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -k "packed or fused or one_launch or batched or rollout or smoke or closed_loop or svmpc" > gpurun_out/pytest_r2_run12.log 2>&1; tail -n 6 gpurun_out/pytest_r2_run12.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout -s KILL 200 python bench.py --no-phi --no-configs --no-cpu-baseline > gpurun_out/bench_r2_run12.json 2> gpurun_out/bench_r2_run12.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r2_run12.json")); print("ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_run12.err").read()[-1500:])
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:svmpc_warp_kernel -s 3 -c 1 -o gpurun_out/fused_r2c -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-phi --no-configs > gpurun_out/ncu_fused_r2c.log 2>&1; tail -n 2 gpurun_out/ncu_fused_r2c.log
