# Collects the round-1 evidence on one B200:  gpurun -- 'bash profiles/r1_collect.sh'
set -x
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench_b.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:svmpc_instance_kernel -s 4 -c 1 -o gpurun_out/fused_r8 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused8.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:noise_normal_kernel -s 1 -c 1 -o gpurun_out/noise_r1 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_noise1.log 2>&1
tail -c 600 gpurun_out/bench_r1_b.json; tail -c 300 gpurun_out/bench_r1_ref.json
