# round 2, second hardware run (one GPU): 3 vs 2 S/P buffers in phi_tc_kernel (A/B libraries), ncu of the
# second-generation instance kernel and of the phi kernel
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core or large_phi or exact_median" > gpurun_out/pytest_phi_sb3.log 2>&1; tail -n 4 gpurun_out/pytest_phi_sb3.log
for v in sb3 sb2; do
  if [ $v = sb2 ]; then export DUST_B200_LIB=$PWD/dust_b200/libdust_b200_sb2.so; else unset DUST_B200_LIB; fi
  timeout -s KILL 200 python bench_phi.py --steps 10 --warmup 3 > gpurun_out/bench_phi_r2_$v.json 2> gpurun_out/bench_phi_r2_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_phi_r2_$v.json")); r = d["roofline"]
    print("$v", "ms_phi", d["ms_phi"], "with median", d["ms_phi_with_median"], "frac", r["frac"], "vs cublas", r["frac_vs_inrun_cublas_tf32"], "err", d.get("rel_err_vs_float64_rows"), d["clocks"])
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_phi_r2_$v.err").read()[-1500:])
PY
done
unset DUST_B200_LIB
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:svmpc_warp_kernel -s 3 -c 1 -o gpurun_out/fused_r2a -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-phi --no-configs > gpurun_out/ncu_fused_r2a.log 2>&1; tail -n 2 gpurun_out/ncu_fused_r2a.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 2 -c 1 -o gpurun_out/phi_r2a -f python bench_phi.py --steps 1 --warmup 1 > gpurun_out/ncu_phi_r2a.log 2>&1; tail -n 2 gpurun_out/ncu_phi_r2a.log
ls -la gpurun_out/*.ncu-rep | tail -3
