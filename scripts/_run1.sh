timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "transforms_vs_fp64_spec or slab_plans or tensor_core_layer_gates" 2>&1 | tail -15
