# Second collection of round 1 (after the persistent phi split, the few-instance paths and the sharded rollout):
#   gpurun --timeout 900 -- 'bash profiles/r1_collect2.sh'
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r1c.log 2>&1; tail -n 15 gpurun_out/pytest_gpu_r1c.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; tail -c 700 gpurun_out/bench_r1_c.json
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref_c.json 2> gpurun_out/bench_r1_ref_c.err
timeout -s KILL 100 python bench_phi.py --steps 5 --warmup 2 --check > gpurun_out/bench_phi_r1_c.json 2> gpurun_out/bench_phi_r1_c.err; tail -c 400 gpurun_out/bench_phi_r1_c.json
for w in 2 4 8; do timeout -s KILL 100 python bench_phi.py --steps 5 --warmup 2 --check --emulate-world $w > gpurun_out/bench_phi_r1_c_emu$w.json 2> gpurun_out/bench_phi_r1_c_emu$w.err; done
timeout -s KILL 200 python bench_configs.py --steps 10 --warmup 3 > gpurun_out/bench_configs_r1_c.json 2> gpurun_out/bench_configs_r1_c.err; cut -c1-420 gpurun_out/bench_configs_r1_c.json
timeout -s KILL 100 python bench_configs.py --configs dual_stress --steps 10 --warmup 3 --alpha 1e-7 > gpurun_out/bench_configs_r1_c_dense.json 2>> gpurun_out/bench_configs_r1_c.err
timeout -s KILL 100 python bench_configs.py --configs dual_stress --steps 10 --warmup 3 --emulate-world 8 > gpurun_out/bench_configs_r1_c_emu8.json 2>> gpurun_out/bench_configs_r1_c.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"phi_tc_kernel|median_tc_kernel" -c 2 -o gpurun_out/tc_r4 -f python bench_phi.py --steps 1 --warmup 0 > gpurun_out/ncu_tc4.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:rollout_cost_kernel -c 1 -o gpurun_out/rollout_r1 -f python bench_configs.py --configs dual_stress --steps 1 --warmup 0 > gpurun_out/ncu_rollout1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
