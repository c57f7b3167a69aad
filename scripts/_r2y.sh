#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "fused_forward_analysis" -s 2>&1 | grep -v "^$" | tail -14
timeout 300 python scripts/block_variants.py 10 2>&1 | grep "fwd"
timeout 900 python bench.py --no-cpu-baseline --no-incumbent --no-other-precision --no-layer-times > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['first_loss'], d['final_loss'])
for k,v in list(d['kernels'].items())[:10]: print(f"{k:32s} {v['launches']:4d} {v['ms_total']/d['steps']:.3f} ms/step")
PY
