timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "stage_split or sharded_projection or slab_plans or vs_oracle" 2>&1 | tail -15
