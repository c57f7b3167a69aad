#!/bin/bash
mkdir -p gpurun_out
echo "== default build (GENG g' formed in step 1)"; timeout 300 python scripts/block_variants.py 10 2>&1 | grep "bwdz\|bwd last\|bwd first"
echo "== ring3 build (first block: three-stage ring)"; PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.ring3.so timeout 300 python scripts/block_variants.py 10 2>&1 | grep "bwdz\|bwd last\|bwd first"
PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.ring3.so timeout 600 python -m pytest tests -x -q -m gpu -k "fused_block_backward or first_block" 2>&1 | tail -3
for v in default mix4 mix5; do
  echo "== $v"
  if [ $v = default ]; then unset PDEPHI_LIB; else export PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.$v.so; fi
  timeout 200 python scripts/k2_bench.py 3 2>&1 | tail -1
done
unset PDEPHI_LIB
timeout 600 python -m pytest tests -x -q -m gpu -k "fused_block_backward or headline or last_block or transfer_function" 2>&1 | tail -3
