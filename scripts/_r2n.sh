#!/bin/bash
mkdir -p gpurun_out
echo "== default (4 loader warps, 672 threads)"; timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2n_block_variants_default.log
echo "== ld3 (3 loader warps, 640 threads)"; PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.ld3.so timeout 300 python scripts/block_variants.py 10 2>&1 | tee gpurun_out/r2n_block_variants_ld3.log
