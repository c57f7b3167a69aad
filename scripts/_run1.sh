timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "transfer_function or rational_layer or golden" 2>&1 | tail -4
timeout 300 python scripts/layer_bench.py tf32 10 --gen 2>&1 | tail -3 | cut -c1-420
