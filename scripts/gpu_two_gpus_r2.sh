#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2_bench_n2.json 2> gpurun_out/r2_n2_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2_n2_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_n2_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d.get('slab'),indent=1)); print(json.dumps(d.get('cfg3_dp'),indent=1)); print(d.get('multi_gpu_blocks'))
PY
timeout 600 python -m pytest tests -q -m gpu -k "slab_parallel_two_gpus" > gpurun_out/r2_n2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_n2_pytest2.log
