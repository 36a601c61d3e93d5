for g in 64 128; do for p in 125 1000; do
QK_SIM_THREADS=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --points $p > gpurun_out/b_g$g.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_g$g.json')); print('G $g N $p', d['ms_per_step'], d['stage_ms'])"
done; done
