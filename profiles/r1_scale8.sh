# 8-GPU scaling of both benches (one box, NVSwitch).  Run: gpurun --gpus 8 -- 'bash profiles/r1_scale8.sh'
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_bench_n$n.json 2> gpurun_out/scale_bench_n$n.err
  timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench_phi.py --steps 5 --warmup 2 > gpurun_out/scale_phi_n$n.json 2> gpurun_out/scale_phi_n$n.err
done
timeout -s KILL 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_bench_n1.json 2> gpurun_out/scale_bench_n1.err
timeout -s KILL 240 python bench_phi.py --steps 5 --warmup 2 > gpurun_out/scale_phi_n1.json 2> gpurun_out/scale_phi_n1.err
tail -n1 gpurun_out/scale_bench_n*.json gpurun_out/scale_phi_n*.json | cut -c1-300
