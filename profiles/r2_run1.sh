# round 2, first hardware run (one GPU): the new parity tests, the whole GPU suite, the new bench line
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl
timeout -s KILL 900 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/pytest_gpu_r2_run1.log 2>&1; tail -n 25 gpurun_out/pytest_gpu_r2_run1.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2_run1.log 2>&1; tail -n 3 gpurun_out/smoke_r2_run1.log
timeout -s KILL 400 python bench.py > gpurun_out/bench_r2_run1.json 2> gpurun_out/bench_r2_run1.err; tail -c 3000 gpurun_out/bench_r2_run1.json; tail -n 5 gpurun_out/bench_r2_run1.err
