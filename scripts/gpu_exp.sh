for m in 0 8 12 9 13 15; do QK_GRAM_DEBUG=$m python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_dbg$m.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_dbg$m.json')); print('dbg=$m gram ms', round(d['stage_ms']['gram'],2))"; done
