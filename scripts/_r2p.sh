#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2p_block_variants.log | grep -v "^{"
timeout 1200 python -m pytest tests -x -q -m gpu -k "golden or oracle or rational or transfer or fused_block_backward or headline or deterministic or cfg4" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-incumbent --no-other-precision > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], 'k2 share', d['k2_share_of_step'])
for k,v in list(d['kernels'].items())[:12]: print(f"{k:32s} {v['launches']:4d} {v['ms_total']/d['steps']:.3f} ms/step")
print(d['layer_fwd_bwd_us'])
PY
