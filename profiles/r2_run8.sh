# round 2, eighth hardware run (one GPU): median kernel with the per-thread hit queues, packed [X | score] read in place
# (row stride), the whole GPU suite, the phi bench, ncu --set full of median_tc_kernel and phi_tc_kernel
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -k "not anisotropic" --durations=5 > gpurun_out/pytest_gpu_r2_run8.log 2>&1; tail -n 12 gpurun_out/pytest_gpu_r2_run8.log
timeout -s KILL 300 python bench_phi.py --steps 10 --warmup 3 > gpurun_out/bench_phi_r2_run8.json 2> gpurun_out/bench_phi_r2_run8.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_phi_r2_run8.json") if l.startswith("{")][-1])
    print("phi", d["ms_phi"], d["ms_phi_with_median"], d["roofline"]["frac"], d.get("rel_err_vs_float64_rows"), d.get("median"), d.get("clocks"))
    print("kernels", {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_phi_r2_run8.err").read()[-2500:])
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:median_tc_kernel -s 2 -c 1 -o gpurun_out/median_r2c -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_median_r2c.log 2>&1; tail -n 2 gpurun_out/ncu_median_r2c.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 4 -c 2 -o gpurun_out/phi_r2c -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_phi_r2c.log 2>&1; tail -n 2 gpurun_out/ncu_phi_r2c.log
