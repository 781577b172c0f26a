# round 2, two GPUs: the bench line under torchrun (phi rows over NCCL with in-run parity, dual-stress draws over NCCL)
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r2_n2.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    p = d["phi"]; print("phi", {k: p.get(k) for k in ("ms_phi", "ms_phi_with_median", "strong_scaling", "strong_scaling_with_median", "max_rel_diff_vs_single_gpu", "bandwidth_bits_equal_to_single_gpu", "rel_err_vs_float64_rows", "single_gpu_same_process")})
    print("phi kernels", p["kernels_ms"])
    c = d["configs"]["dual_stress"]; print("stress", {k: c.get(k) for k in ("device_ms_per_dual_step", "single_gpu_same_process_ms", "strong_scaling", "max_rel_diff_vs_single_gpu", "collectives")})
    print("stress kernels", c["kernel_ms_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_n2.err").read()[-3000:])
PY
for g in packed separate; do
  timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
      bench_phi.py --steps 10 --warmup 3 --gather $g > gpurun_out/bench_phi_r2_n2_$g.json 2> gpurun_out/bench_phi_r2_n2_$g.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_phi_r2_n2_$g.json')); print('$g', d['ms_phi'], d['ms_phi_with_median'], d.get('strong_scaling'), d.get('strong_scaling_with_median'), d.get('max_rel_diff_vs_single_gpu'))" || tail -n 20 gpurun_out/bench_phi_r2_n2_$g.err
done
