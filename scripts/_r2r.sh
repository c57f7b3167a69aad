#!/bin/bash
mkdir -p gpurun_out
for v in default za4 za2; do
  echo "== $v"
  if [ $v = default ]; then unset PDEPHI_LIB; else export PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.$v.so; fi
  timeout 300 python scripts/block_variants.py 10 2>&1 | grep "bwdz"
done
