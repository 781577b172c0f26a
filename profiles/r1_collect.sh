set -x
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
timeout -s KILL 300 python bench_phi.py --check > gpurun_out/bench_phi_r1.json 2> gpurun_out/bench_phi_r1.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench_b.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:svmpc_instance_kernel -s 4 -c 1 -o gpurun_out/fused_r5 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused5.log 2>&1
tail -2 gpurun_out/bench_r1_b.json; tail -1 gpurun_out/bench_r1_ref.json; tail -3 gpurun_out/bench_r1_b.err
