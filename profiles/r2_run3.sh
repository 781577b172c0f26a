# round 2, third hardware run (one GPU): instance kernel after the per-tile clean-up (host-side coefficients, one buffer per
# warp, templated fold), parity subset, A/B against the first-generation kernel, ncu capture
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -k "packed or fused or one_launch or batched or rollout or smoke or adjoint or mpf" > gpurun_out/pytest_r2_run3.log 2>&1; tail -n 4 gpurun_out/pytest_r2_run3.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
for v in v1 v2; do
  if [ $v = v1 ]; then export DUST_B200_FUSED_V1=1; else unset DUST_B200_FUSED_V1; fi
  timeout -s KILL 200 python bench.py --no-phi --no-configs --no-cpu-baseline > gpurun_out/bench_r2_run3_$v.json 2> gpurun_out/bench_r2_run3_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2_run3_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_r2_run3_$v.err").read()[-1500:])
PY
done
unset DUST_B200_FUSED_V1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:svmpc_warp_kernel -s 3 -c 1 -o gpurun_out/fused_r2b -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-phi --no-configs > gpurun_out/ncu_fused_r2b.log 2>&1; tail -n 2 gpurun_out/ncu_fused_r2b.log
