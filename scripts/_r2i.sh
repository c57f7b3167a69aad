#!/bin/bash
# round 2, session 3: state check of HEAD -- full GPU test suite, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2i_pytest.log
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2i_bench.err
head -c 1200 gpurun_out/r2i_bench.json
