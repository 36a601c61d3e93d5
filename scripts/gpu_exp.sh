python scripts/gpu_c4.py 2>&1 | grep "^G="
QK_SIM_THREADS=128 python scripts/gpu_c4.py 2>&1 | grep "^G="
