# round 2, first hardware run (one GPU): new parity tests + whole GPU suite, the new bench line, and the
# second-generation instance kernel against the first (A/B through DUST_B200_FUSED_V1 / DUST_B200_NBUF)
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl
timeout -s KILL 1200 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/pytest_gpu_r2_run1.log 2>&1; tail -n 30 gpurun_out/pytest_gpu_r2_run1.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2_run1.log 2>&1; tail -n 3 gpurun_out/smoke_r2_run1.log
for v in v1 4 5 6; do
  if [ $v = v1 ]; then export DUST_B200_FUSED_V1=1; else unset DUST_B200_FUSED_V1; export DUST_B200_NBUF=$v; fi
  timeout -s KILL 200 python bench.py --no-phi --no-configs --no-cpu-baseline > gpurun_out/bench_r2_ab_$v.json 2> gpurun_out/bench_r2_ab_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2_ab_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e:
    print("$v failed", e)
PY
done
unset DUST_B200_FUSED_V1 DUST_B200_NBUF
timeout -s KILL 500 python bench.py > gpurun_out/bench_r2_run1.json 2> gpurun_out/bench_r2_run1.err; tail -c 1500 gpurun_out/bench_r2_run1.json; tail -n 5 gpurun_out/bench_r2_run1.err
timeout -s KILL 300 compute-sanitizer --tool memcheck --launch-timeout 0 python -m pytest tests/test_gpu_parity.py -q -x -k "packed_pendulum or fused_instance_kernel" > gpurun_out/sanitize_v2_memcheck.log 2>&1; tail -n 6 gpurun_out/sanitize_v2_memcheck.log
timeout -s KILL 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_api.py -q -x -k "one_launch" > gpurun_out/sanitize_v2_racecheck.log 2>&1; tail -n 6 gpurun_out/sanitize_v2_racecheck.log
