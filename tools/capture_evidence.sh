set -x
mkdir -p gpurun_out
PARITY_LOG=gpurun_out/parity_report.json timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
timeout 400 ncu --set full --import-source on --clock-control none -k regex:emit_small -c 1 --launch-skip 4 -o gpurun_out/r2_emit_small_full -f python bench.py --steps 4 --warmup 3 --no-cpu --configs none --e2e-steps 1 > gpurun_out/ncu_f.log 2>&1; echo "full rc=$?"
python tools/ncu_traffic.py gpurun_out/r2_emit_small_full.ncu-rep 16777216 > gpurun_out/ncu_traffic.log 2>&1; cp profiles/emit_traffic.json gpurun_out/emit_traffic.json
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --configs none --e2e-steps 1 > gpurun_out/ncu_l.log 2>&1; echo "launches rc=$?"
tail -3 gpurun_out/pytest_gpu.log
