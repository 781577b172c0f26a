# round 2, twenty-fifth hardware run (one GPU): svmpc_quad_kernel (two pairs per lane, 16 warps/SM) against svmpc_warp_kernel
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_bench_contract.py -q -x -k "packed or fused or one_launch or batched or rollout or smoke or closed_loop or svmpc or bench" > gpurun_out/pytest_r2_run25.log 2>&1; tail -n 5 gpurun_out/pytest_r2_run25.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
for v in quad pair; do
  if [ $v = pair ]; then export DUST_B200_NO_QUAD=1; else unset DUST_B200_NO_QUAD; fi
  timeout -s KILL 200 python bench.py --no-phi --no-configs --no-cpu-baseline > gpurun_out/bench_r2_run25_$v.json 2> gpurun_out/bench_r2_run25_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2_run25_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_r2_run25_$v.err").read()[-1500:])
PY
done
unset DUST_B200_NO_QUAD
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:svmpc_quad_kernel -s 3 -c 1 -o gpurun_out/fused_r2e -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-phi --no-configs > gpurun_out/ncu_fused_r2e.log 2>&1; tail -n 1 gpurun_out/ncu_fused_r2e.log
