set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err; cut -c1-600 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.err; cut -c1-400 gpurun_out/bench_ours.json
python bench.py --workload c3_g0.1 --no-cpu-baseline > gpurun_out/bench_ours_g01.json 2>> gpurun_out/bench_ours.err; cut -c1-300 gpurun_out/bench_ours_g01.json
