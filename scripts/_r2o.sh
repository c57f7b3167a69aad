#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2o_block_variants.log | grep -v "^{"
timeout 600 python -m pytest tests -x -q -m gpu -k "fused_block_backward or headline" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-incumbent --no-layer-times --no-other-precision > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in list(d['kernels'].items())[:8]: print(f"{k:32s} {v['launches']:4d} {v['ms_total']/d['steps']:.3f} ms/step")
PY
