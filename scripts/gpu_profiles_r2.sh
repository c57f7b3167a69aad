#!/bin/bash
# round 2 profiles: launch list of the bench step and a --set full capture of every kernel of one fwd+bwd pass of a 3-block
# model (first / middle / last block variants) on the final build; the .ncu-rep is summarised on the box (gpurun_out is
# limited to 64 MiB) and only the source-level report of the dominant kernel is brought back
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-times --no-other-precision --no-incumbent"
$CMD > gpurun_out/r2_prof_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/r2_launches_bench_tf32.csv $CMD > gpurun_out/r2_prof_ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_bench_tf32.csv
python scripts/profile_model_blocks.py tf32 2 > gpurun_out/r2_prof_plain_blocks.log 2>&1 && \
ncu --set full --clock-control none -k regex:"block_tail_fwd_kernel|block_tail_bwd_kernel|gen_plane_analysis|rational_dpq_pipe|gen_axis_synthesis|analysis_zy_tc_kernel|synthesis_zy_tc_kernel|rational_mix_kernel|mix_from_R|rational_g_kernel|analysis_x_tc|synthesis_x_tc" -s 45 -c 45 -f -o /tmp/r2_prof_blocks \
    python scripts/profile_model_blocks.py tf32 2 > gpurun_out/r2_prof_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_prof_ncu_full.log
python scripts/ncu_summary.py /tmp/r2_prof_blocks.ncu-rep gpurun_out/r2_ncu_full_kernels.csv --every; wc -l gpurun_out/r2_ncu_full_kernels.csv
du -sh gpurun_out
