#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2l_pytest.log
grep -h "headline tf32\|fused block backward\|fused vs unfused" gpurun_out/r2l_pytest.log | head
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2l_bench.err
head -c 400 gpurun_out/r2l_bench.json
