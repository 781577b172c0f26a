# round 2, twenty-third hardware run (one GPU): compute-sanitizer over the kernels this round rewrote
mkdir -p gpurun_out
SEL='partitions and (8192-40-rows1 or 8192-40-rows2 or 8192-24) or packed_buffer or peer_gather or tensor_core_median_vs_radix_select and low_dim'
for tool in memcheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --launch-timeout 0 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/sanitize_r2_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanitize_r2_$tool.log | tail -n 4
done
timeout -s KILL 600 compute-sanitizer --tool memcheck --launch-timeout 0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -k "packed_pendulum or fused_instance_kernel or one_launch" > gpurun_out/sanitize_r2_memcheck_instance.log 2>&1
echo "== memcheck instance kernel"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_r2_memcheck_instance.log | tail -n 3
