set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err
python scripts/layer_bench.py tf32 10 > gpurun_out/layer_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"block_tail|analysis_zy_tc|synthesis_zy_tc|rational_mix_kernel|rational_g_kernel|rational_dpq|mix_from_R" -s 30 -c 14 -o gpurun_out/prof_s2 -f python scripts/profile_layer_block.py > gpurun_out/ncu_s2.log 2>&1
tail -3 gpurun_out/pytest3.log
