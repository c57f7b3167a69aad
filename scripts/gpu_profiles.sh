# Round-1 (second half) profile captures: run on the GPU box through gpurun; outputs land in gpurun_out/.
set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-times --no-other-precision > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 600 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-times --no-other-precision > gpurun_out/ncu_launch_b.log 2>&1
python scripts/profile_layer_block.py tf32 4 || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:"block_tail_fwd|block_tail_bwd_kernel|analysis_zy_tc_kernel|synthesis_zy_tc_kernel|analysis_x_tc|synthesis_x_tc|rational_mix_kernel|rational_g_kernel|rational_dpq|mix_from_R|projection" \
    -s 40 -c 20 -o gpurun_out/prof_r1b -f python scripts/profile_layer_block.py tf32 4 > gpurun_out/ncu_full_b.log 2>&1
python scripts/profile_layer_block.py tf32x3 4 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"tc3" -s 12 -c 8 -o gpurun_out/prof_r1b_tc3 -f \
    python scripts/profile_layer_block.py tf32x3 4 > gpurun_out/ncu_full_b3.log 2>&1
tail -2 gpurun_out/ncu_full_b.log gpurun_out/ncu_full_b3.log
