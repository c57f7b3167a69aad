#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2g_block_variants.log
timeout 900 python -m pytest tests -x -q -m gpu -k "block or lift or head or fused or cfg3 or cfg2" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2g_pytest.log
