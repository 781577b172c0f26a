# round 2, thirtieth hardware run (one GPU): e2e with the next step's noise drawn on a side stream (A/B)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_round2.py tests/test_bench_contract.py -q -x -k "prefetch or bench" 2>&1 | tail -n 3
for v in prefetch serial; do
  if [ $v = serial ]; then export DUST_BENCH_NO_PREFETCH=1; else unset DUST_BENCH_NO_PREFETCH; fi
  timeout -s KILL 200 python bench.py --no-phi --no-configs --no-cpu-baseline > gpurun_out/bench_r2_run30_$v.json 2> gpurun_out/bench_r2_run30_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2_run30_$v.json")); print("$v", "ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_r2_run30_$v.err").read()[-1500:])
PY
done
