set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qk_gram_dmma -s 3 -c 1 -o gpurun_out/prof_gram_r01g $CMD > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log
