set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01z.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qk_sim_kernel -s 3 -c 1 -o gpurun_out/prof_sim_r01z $CMD > gpurun_out/ncu_sim.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qk_gram_dmma -s 3 -c 1 -o gpurun_out/prof_gram_r01z $CMD > gpurun_out/ncu_gram.log 2>&1
tail -3 gpurun_out/ncu_sim.log gpurun_out/ncu_gram.log gpurun_out/ncu_launch.log
ls -la gpurun_out
