for v in 4x1 4x3 2x2 8x1; do QKMPS_LIB=$PWD/qml-cutensornet_b200/qkmps/libqkmps_t$v.so python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_t$v.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_t$v.json')); print('tile $v gram ms', round(d['stage_ms']['gram'],2))"; done
tail -3 gpurun_out/b.err
