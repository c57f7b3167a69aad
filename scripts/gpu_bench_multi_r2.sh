#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_multi_bench_n$N.err; echo "bench n$N rc=$?"
tail -c 800 gpurun_out/r2_multi_bench_n$N.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d.get('slab'))); print(json.dumps(d.get('cfg3_dp'))); print(d.get('multi_gpu_blocks'))
PY
