#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fused_bwd_debug.py > gpurun_out/r2j_fused_bwd.log 2>&1; echo "debug rc=$?"
tail -60 gpurun_out/r2j_fused_bwd.log
