#!/bin/bash
echo "== default (GENG: step 1 forms g, 3 loaders)"; timeout 300 python scripts/block_variants.py 10 2>&1 | grep "last"
echo "== gl (GENG: loaders generate, 4 loaders)"; PDEPHI_LIB=$PWD/pde_fluid_phi_b200/lib/libpdephi_b200.gl.so timeout 300 python scripts/block_variants.py 10 2>&1 | grep "last"
timeout 200 python scripts/k2_bench.py 3 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu -k "fused_block_backward or last_block or headline" 2>&1 | tail -2
