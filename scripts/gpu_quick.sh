set -x
python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_default.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_default.json')); print('default', d['ms_per_step'], d['stage_ms'], d['value'], d['stages']['gram']['frac'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --points 125 > gpurun_out/b_p125.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_p125.json')); print('N=125', d['ms_per_step'], d['stage_ms'], d['value'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c3_g0.1 > gpurun_out/b_g01.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_g01.json')); print('gamma0.1', d['ms_per_step'], d['stage_ms'], d['value'])"
tail -5 gpurun_out/b.err
