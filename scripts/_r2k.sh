#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fused_bwd_debug.py > gpurun_out/r2j_fused_bwd.log 2>&1; echo "debug rc=$?"
tail -22 gpurun_out/r2j_fused_bwd.log
timeout 1500 python -m pytest tests -x -q -m gpu -k "headline or block or lift or head or cfg2 or cfg3 or fused or deterministic or rollout" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2k_pytest.log
