for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/slab_bench.py tf32 2>&1 | grep -E "cfg-5|Error|error" | tail -3
done
