#!/bin/bash
# round 2: --set full capture of the unfused block kernels as the fp32-grade (`auto`) route runs them (packed derivative)
mkdir -p gpurun_out
python scripts/profile_model_blocks.py auto 2 > gpurun_out/r2_prof_plain_blocks_auto.log 2>&1 && \
ncu --set full --clock-control none -k regex:"block_tail_fwd_kernel|block_tail_bwd_kernel|analysis_zy_tc3|synthesis_zy_tc3" -s 14 -c 20 -f -o /tmp/r2_prof_auto \
    python scripts/profile_model_blocks.py auto 2 > gpurun_out/r2_prof_ncu_full_auto.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_prof_ncu_full_auto.log
python scripts/ncu_summary.py /tmp/r2_prof_auto.ncu-rep gpurun_out/r2_ncu_full_auto_block_kernels.csv --every; wc -l gpurun_out/r2_ncu_full_auto_block_kernels.csv
