python scripts/profile_layer_block.py tf32 3 || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:"block_tail_fwd|block_tail_bwd_kernel|analysis_zy_tc_kernel|synthesis_zy_tc_kernel|gen_axis_synthesis|analysis_x_tc|synthesis_x_tc|rational_mix_kernel" \
    -s 11 -c 24 -o gpurun_out/prof_r1d -f python scripts/profile_layer_block.py tf32 3 > gpurun_out/ncu_full_d2.log 2>&1
tail -2 gpurun_out/ncu_full_d2.log
