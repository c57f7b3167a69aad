timeout 900 python -m pytest tests/test_configs_gpu.py -x -q -k cfg5 2>&1 | grep -E "Error|rel\(|passed|failed|^E" | head -20
timeout 300 python scripts/gen_debug.py small 2>&1 | tail -24
