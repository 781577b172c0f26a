# round 2, eighteenth hardware run (one GPU): GEMM2's P_lo V term as kind::f16 on bf16 copies, S ring of three
mkdir -p gpurun_out
for sw in 0 1; do
  if [ $sw = 1 ]; then export DUST_B200_TC_PLO_SWAP=1; else unset DUST_B200_TC_PLO_SWAP; fi
  echo "--- P_lo packing swap = $sw"
  timeout -s KILL 300 python -m pytest tests/test_gpu_round2.py -q -x -k "partitions" 2>&1 | tail -n 4
  timeout -s KILL 120 python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/parity_table.jsonl")]
for r in rows[-14:]:
    if "phi partitions" in r["case"] or "operand modes" in r["case"]: print(r)
PY
done
unset DUST_B200_TC_PLO_SWAP
timeout -s KILL 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "(phi or median or svgd) and not anisotropic" --durations=3 > gpurun_out/pytest_gpu_r2_run18.log 2>&1; tail -n 6 gpurun_out/pytest_gpu_r2_run18.log
for w in 1 8; do
  timeout -s KILL 300 python bench_phi.py --steps 10 --warmup 3 --emulate-world $w > gpurun_out/bench_phi_r2_run18_w$w.json 2> gpurun_out/bench_phi_r2_run18_w$w.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_phi_r2_run18_w$w.json") if l.startswith("{")][-1])
    print("world $w", "ms_phi", d["ms_phi"], "with median", d["ms_phi_with_median"], "frac", d.get("roofline", {}).get("frac"), d.get("rel_err_vs_float64_rows"), d.get("clocks"))
    print("   kernels", {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_phi_r2_run18_w$w.err").read()[-2500:])
PY
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 4 -c 2 -o gpurun_out/phi_r2f -f python bench_phi.py --steps 2 --warmup 2 --no-checks > gpurun_out/ncu_phi_r2f.log 2>&1; tail -n 2 gpurun_out/ncu_phi_r2f.log
