#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_headline_gate_gpu.py -x -q -m gpu -k "transforms_vs_fp64_spec or per_axis or cfg2_shape or headline or default_precision" > gpurun_out/r2d_x3_tests.log 2>&1; echo "x3 tests rc=$?"
tail -5 gpurun_out/r2d_x3_tests.log
for p in auto tf32; do
  timeout 300 python scripts/layer_bench.py $p 10 --gen --gen-only > gpurun_out/r2d_layers_$p.log 2>&1; echo "layers $p rc=$?"
  cat gpurun_out/r2d_layers_$p.log
done
