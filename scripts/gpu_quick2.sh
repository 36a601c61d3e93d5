python -m pytest tests -q -m gpu 2>&1 | tail -3
python scripts/gpu_configs.py 2>&1 | grep -E "^C|^run|Error"
