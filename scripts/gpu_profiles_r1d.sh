# Round-1d profile captures (fused spectral forward, reordered loaders): run on the GPU box through gpurun.
set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-times --no-other-precision > gpurun_out/bench_for_ncu_d.json 2> gpurun_out/bench_for_ncu_d.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 440 --csv --log-file gpurun_out/launches_r1d.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-times --no-other-precision > gpurun_out/ncu_launch_d.log 2>&1
python scripts/profile_layer_block.py tf32 4 || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:"block_tail_fwd|block_tail_bwd_kernel|analysis_zy_tc_kernel|synthesis_zy_tc_kernel|gen_axis_synthesis|analysis_x_tc|synthesis_x_tc|rational_mix_kernel" \
    -s 11 -c 24 -o gpurun_out/prof_r1d -f python scripts/profile_layer_block.py tf32 4 > gpurun_out/ncu_full_d2.log 2>&1
tail -2 gpurun_out/ncu_full_d2.log
wc -l gpurun_out/launches_r1d.csv
