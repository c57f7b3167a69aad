set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "slab_parallel_two_gpus" 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_bench.py tf32 2>&1 | grep -v "^W\|^\[W" | tail -8
