set -x
python -m pytest tests -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 --cpu-budget 8 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --steps 3 --warmup 3 --workload c3_g0.1 --no-cpu-baseline > gpurun_out/bench_c3_g01.json 2>> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3_g01.json
