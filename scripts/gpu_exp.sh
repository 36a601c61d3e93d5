python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_default.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_default.json')); print('default', d['ms_per_step'], d['stage_ms'])"
for cfg in "128 16" "64 8" "32 8" "64 16" "32 16"; do set -- $cfg; QK_SIM_THREADS=$1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c3_g0.1 --chi $2 > gpurun_out/b_x.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_x.json')); print('gamma0.1 G=$1 chi=$2', d['ms_per_step'], d['stage_ms'])"; done
tail -3 gpurun_out/b.err
