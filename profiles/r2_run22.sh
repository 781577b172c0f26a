# round 2, twenty-second hardware run (one GPU): relaxed wait of the flush warps; launch list of the bench command; final ncu captures
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "(phi or median or svgd or peer) and not 65536" > gpurun_out/pytest_gpu_r2_run22.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_r2_run22.log
for w in 1 8; do
  timeout -s KILL 300 python bench_phi.py --steps 10 --warmup 3 --emulate-world $w > gpurun_out/bench_phi_r2_run22_w$w.json 2> gpurun_out/bench_phi_r2_run22_w$w.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_phi_r2_run22_w$w.json") if l.startswith("{")][-1])
    print("world $w", "ms_phi", d["ms_phi"], "with median", d["ms_phi_with_median"], "frac", d.get("roofline", {}).get("frac"), d.get("rel_err_vs_float64_rows"), d.get("clocks"))
    print("   kernels", {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_phi_r2_run22_w$w.err").read()[-2500:])
PY
done
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --phi-steps 1 --config-steps 2 --no-cpu-baseline > gpurun_out/launches_r2_bench.log 2>&1; tail -n 1 gpurun_out/launches_r2_bench.log | cut -c1-200
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 4 -c 2 -o gpurun_out/phi_r2i -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_phi_r2i.log 2>&1; tail -n 1 gpurun_out/ncu_phi_r2i.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:svmpc_warp_kernel -s 3 -c 1 -o gpurun_out/fused_r2d -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-phi --no-configs > gpurun_out/ncu_fused_r2d.log 2>&1; tail -n 1 gpurun_out/ncu_fused_r2d.log
