set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-times --no-other-precision > gpurun_out/ncu_launch_b.log 2>&1
wc -l gpurun_out/launches_r1b.csv
