# compute-sanitizer passes over the kernels of round 1 (memcheck everywhere, racecheck on the
# shared-memory / mbarrier heavy ones).  gpurun -- 'bash profiles/sanitize.sh'
mkdir -p gpurun_out
SEL='fused_instance_kernel or one_launch or adjoint_matches or adjoint_batched or normal_2048 or low_dim or roll_strategy or batched_instances or mpf_against'
timeout -s KILL 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
timeout -s KILL 500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "fused_instance_kernel or one_launch or adjoint_batched" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log | tail -12
