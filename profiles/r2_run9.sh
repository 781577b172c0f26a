# round 2, ninth hardware run (one GPU): row tile of GEMM1 in TMEM (TS form) and the column-chunk partition of phi_tc_kernel
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "phi or median or svgd" --durations=5 > gpurun_out/pytest_gpu_r2_run9.log 2>&1; tail -n 12 gpurun_out/pytest_gpu_r2_run9.log
for mode in tmem smem; do
  if [ $mode = smem ]; then export DUST_B200_TC_A_SMEM=1; else unset DUST_B200_TC_A_SMEM; fi
  for w in 1 8; do
    timeout -s KILL 300 python bench_phi.py --steps 10 --warmup 3 --emulate-world $w --no-checks > gpurun_out/bench_phi_r2_run9_${mode}_w$w.json 2> gpurun_out/bench_phi_r2_run9_${mode}_w$w.err
    python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_phi_r2_run9_${mode}_w$w.json") if l.startswith("{")][-1])
    print("$mode", "world $w", "ms_phi", d["ms_phi"], "with median", d["ms_phi_with_median"], "frac", d.get("roofline", {}).get("frac"), d.get("clocks"))
    print("   kernels", {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_phi_r2_run9_${mode}_w$w.err").read()[-2500:])
PY
  done
done
unset DUST_B200_TC_A_SMEM
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 4 -c 2 -o gpurun_out/phi_r2d -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_phi_r2d.log 2>&1; tail -n 2 gpurun_out/ncu_phi_r2d.log
