for lib in libqkmps_t2.so libqkmps.so; do
QKMPS_LIB=$PWD/qml-cutensornet_b200/qkmps/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_x.json 2>> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b_x.json')); print('$lib', round(d['ms_per_step'],2), d['stage_ms'])"
done
QKMPS_LIB=$PWD/qml-cutensornet_b200/qkmps/libqkmps_t2.so python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gram or Gram or kernel" 2>&1 | tail -2
tail -3 gpurun_out/b.err
