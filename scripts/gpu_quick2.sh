python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_default.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_default.json')); print('default', d['ms_per_step'], d['stage_ms'], d['value'], d['stages']['gram']['frac'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c3_g0.1 > gpurun_out/b_g01.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_g01.json')); print('gamma0.1', d['ms_per_step'], d['stage_ms'], d['value'])"
python scripts/gpu_configs.py 2>&1 | grep -E "^C|^run|Error"
tail -3 gpurun_out/b.err
