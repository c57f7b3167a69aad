timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "block" 2>&1 | tail -12
