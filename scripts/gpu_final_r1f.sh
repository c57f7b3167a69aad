# Round-1f: the incumbent (stock PyTorch / cuFFT on the same B200) beside the library, and a --set full capture of every
# variant of the block kernels on the final build.  Run on the GPU box through gpurun.
set -x
timeout 220 python scripts/incumbent_torch_cufft.py > gpurun_out/incumbent.log 2> gpurun_out/incumbent.err
echo "incumbent rc=$?"
timeout 90 python scripts/profile_model_blocks.py tf32 2 || exit 1
timeout 170 ncu --set full --clock-control none --import-source on -k regex:"block_tail_fwd_kernel|block_tail_bwd_kernel" -s 6 -c 6 -o gpurun_out/prof_r1f -f \
    python scripts/profile_model_blocks.py tf32 2 > gpurun_out/ncu_full_f.log 2>&1
tail -2 gpurun_out/ncu_full_f.log
ls -la gpurun_out/prof_r1f.ncu-rep
