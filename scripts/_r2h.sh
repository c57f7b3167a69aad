#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "block or fused or whole" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2h_pytest.log
timeout 300 python scripts/block_variants.py 10 > gpurun_out/r2h_variants.log 2>&1; echo "variants rc=$?"
grep -v "^{" gpurun_out/r2h_variants.log | tail -14
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['first_loss'], d['final_loss'])
for k,v in list(d['kernels'].items())[:6]: print(k, v)
print(d['other_precision']['value'], d['other_precision']['ms_per_step'])
PY
