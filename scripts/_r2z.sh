#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_fwd_kernel.py > gpurun_out/r2z_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:block_tail_fwd -s 2 -c 1 -f -o gpurun_out/r2z_fwd python scripts/profile_fwd_kernel.py > gpurun_out/r2z_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
