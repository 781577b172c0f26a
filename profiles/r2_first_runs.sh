# Hardware runs owed from round 1 (its GPU budget ran out): run these first in round 2.
#   1 GPU :  gpurun --timeout 600 -- 'bash profiles/r2_first_runs.sh one'
#   2 GPUs:  gpurun --gpus 2 --timeout 300 -- 'bash profiles/r2_first_runs.sh two'
#   8 GPUs:  gpurun --gpus 8 --timeout 300 -- 'bash profiles/r2_first_runs.sh eight'
mkdir -p gpurun_out
case "$1" in
one)
  timeout -s KILL 300 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_r2a.log
  timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; tail -n 3 gpurun_out/smoke_r2a.log
  timeout -s KILL 200 python bench.py > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 400 gpurun_out/bench_r2_a.json
  ;;
two)   # the NCCL path of ShardedRollout (host logic is covered under gloo)
  timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      bench_configs.py --configs dual_stress --steps 10 --warmup 3 > gpurun_out/stress_n2.json 2> gpurun_out/stress_n2.err; tail -c 600 gpurun_out/stress_n2.json
  ;;
eight) # packed vs separate gather of [X | score]
  for g in packed separate; do
    timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        bench_phi.py --steps 10 --warmup 3 --gather $g > gpurun_out/scale_phi_n8_$g.json 2> gpurun_out/scale_phi_n8_$g.err; tail -c 500 gpurun_out/scale_phi_n8_$g.json
  done
  ;;
esac
