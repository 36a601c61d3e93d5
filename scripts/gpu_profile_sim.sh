set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --points 125"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qk_sim_kernel -s 3 -c 1 -o gpurun_out/prof_sim_lat_r01 $CMD > gpurun_out/ncu_sim.log 2>&1
tail -2 gpurun_out/ncu_sim.log
