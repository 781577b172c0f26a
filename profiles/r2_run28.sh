# round 2, twenty-eighth hardware run (one GPU): median_tc_kernel with the row tile in TMEM (TS-form GEMM1)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "(median or phi or svgd) and not anisotropic" --durations=3 > gpurun_out/pytest_gpu_r2_run28.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_r2_run28.log
timeout -s KILL 300 python bench_phi.py --steps 10 --warmup 3 > gpurun_out/bench_phi_r2_run28.json 2> gpurun_out/bench_phi_r2_run28.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_phi_r2_run28.json") if l.startswith("{")][-1])
    print("ms_phi", d["ms_phi"], "with median", d["ms_phi_with_median"], d.get("median"), d.get("clocks"))
    print("   kernels", {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_phi_r2_run28.err").read()[-2500:])
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:median_tc_kernel -s 2 -c 1 -o gpurun_out/median_r2e -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_median_r2e.log 2>&1; tail -n 1 gpurun_out/ncu_median_r2e.log
