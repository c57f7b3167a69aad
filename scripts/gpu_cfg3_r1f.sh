# cfg 3 at its full width: block-route parity test on one GPU, then (with --gpus 2) the data-parallel step with the flat and
# the overlapped gradient all-reduce.
set -x
timeout 100 python -m pytest tests/test_configs_gpu.py -q -x -k "cfg3" 2>&1 | tail -15
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_cfg3_bench.py 2>&1 | tail -12
else
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_cfg3_bench.py 2>&1 | tail -12
fi
