# round 2, thirty-third hardware run (one GPU): captures on the final sources (comment-only changes since run 29) for traffic.json
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_round2.py -q -x -k "partitions or packed_buffer" 2>&1 | tail -n 2
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 4 -c 2 -o gpurun_out/phi_r2l -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_phi_r2l.log 2>&1; tail -n 1 gpurun_out/ncu_phi_r2l.log
