for p in 125 250 500 1000; do for g in 64 128; do
QK_SIM_THREADS=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --points $p > gpurun_out/b_g$g.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_g$g.json')); print('N $p G $g', round(d['ms_per_step'],2), d['stage_ms'])"
done; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c3_g0.1 > gpurun_out/b_g01.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_g01.json')); print('gamma0.1', d['ms_per_step'], d['stage_ms'])"
