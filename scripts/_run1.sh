timeout 300 python scripts/tc_debug.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_comp.json 2> gpurun_out/bench_comp.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_comp.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['first_loss'], d['final_loss'])
for k,v in d['rooflines'].items(): print(k, v['avg_launch_ms'], v['frac'])
PY
