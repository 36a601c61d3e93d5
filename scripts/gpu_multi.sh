set -x
NG=${1:-2}
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU|Error|error" | head -5
for n in $(seq 1 8); do
  if [ $n -le $NG ] && { [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; }; then
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_g$n.err > gpurun_out/bench_g$n.json
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 3 --warmup 3 2>gpurun_out/bench_g$n.err > gpurun_out/bench_g$n.json
    fi
    python -c "
import json
for l in open('gpurun_out/bench_g$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$n', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, round(d['value']), 'e2e', round(d['e2e']['value']))"
  fi
done
