#!/bin/bash
mkdir -p gpurun_out
for v in "" ld3; do
  if [ -z "$v" ]; then unset PDEPHI_LIB; else export PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.$v.so; fi
  echo "== variant ${v:-default}"
  timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2f_block_variants_${v:-default}.log
done
