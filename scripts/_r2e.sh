#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench n2 rc=$?"
tail -c 3000 gpurun_out/r2e_bench_n2.err
timeout 600 python -m pytest tests -q -m gpu -k "slab_parallel_two_gpus or headline" > gpurun_out/r2e_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2e_pytest2.log
