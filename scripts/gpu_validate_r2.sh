#!/bin/bash
# round 2: full GPU suite + default bench line + reference arm on the final build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_validate_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_validate_pytest.log
timeout 900 python bench.py > gpurun_out/r2_validate_bench.json 2> gpurun_out/r2_validate_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_validate_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_validate_bench_reference.json 2> gpurun_out/r2_validate_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/r2_validate_bench_reference.json | head -c 600
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2_validate_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_validate_smoke.log
