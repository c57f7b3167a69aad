timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "pointwise" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-layer-times > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['other_precision'])
for k,v in d['kernels'].items(): print(k, v['launches'], v['ms_total'])
PY
