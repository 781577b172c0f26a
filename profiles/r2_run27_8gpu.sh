# round 2, eight GPUs, fifth collection (final kernels: bf16 P_lo term, S ring of three, one fence per CTA in the push kernel instead of the NCCL all-gather; NCCL gather for comparison): the bench line under torchrun (instances x8; phi rows over NCCL with in-run parity and strong scaling;
timeout -s KILL 300 python -m pytest tests/test_gpu_round2.py -q -x -k "partitions and 4096-37" 2>&1 | tail -n 2
# dual-stress draws over NCCL), then the phi microbench with the packed gather for comparison
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29595 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8e.json 2> gpurun_out/bench_r2_n8e.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2_n8e.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    p = d["phi"]; print("phi", {k: p.get(k) for k in ("gather", "ms_phi", "ms_phi_with_median", "strong_scaling", "strong_scaling_with_median", "max_rel_diff_vs_single_gpu", "bandwidth_bits_equal_to_single_gpu", "rel_err_vs_float64_rows", "single_gpu_same_process")})
    print("phi kernels", {k: round(v, 4) for k, v in p["kernels_ms"].items()}, p["clocks"])
    c = d["configs"]["dual_stress"]; print("stress", {k: c.get(k) for k in ("device_ms_per_dual_step", "single_gpu_same_process_ms", "strong_scaling", "max_rel_diff_vs_single_gpu", "collectives")})
    print("stress kernels", {k: round(v, 4) for k, v in c["kernel_ms_per_step"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_n8e.err").read()[-3000:])
PY
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29596 \
    bench_phi.py --steps 20 --warmup 5 --gather packed > gpurun_out/bench_phi_r2_n8e_packed.json 2> gpurun_out/bench_phi_r2_n8e_packed.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_phi_r2_n8e_packed.json') if l.startswith('{')][-1]); print('packed (NCCL)', d['ms_phi'], d['ms_phi_with_median'], d.get('strong_scaling'), d.get('strong_scaling_with_median'), d.get('max_rel_diff_vs_single_gpu'))" || tail -n 20 gpurun_out/bench_phi_r2_n8e_packed.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29597 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r2_n4e.json 2> gpurun_out/bench_r2_n4e.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2_n4e.json") if l.startswith("{")][-1])
    p = d["phi"]; print("4 GPUs: value", d["value"], "phi", {k: p.get(k) for k in ("gather", "gather_note", "ms_phi", "ms_phi_with_median", "strong_scaling", "strong_scaling_with_median", "max_rel_diff_vs_single_gpu")}, "stress", d["configs"]["dual_stress"].get("strong_scaling"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2_n4e.err").read()[-2000:])
PY
