#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2_block_variants.log | grep -v "^{"
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2v_pytest.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], 'k2 share', d['k2_share_of_step'], d['roofline']['kernel'], d['roofline']['frac'])
for k,v in list(d['kernels'].items())[:6]: print(f"{k:32s} {v['launches']:4d} {v['ms_total']/d['steps']:.3f} ms/step")
PY
