for d in 0 1 2 3; do echo "== dbg $d"; PDEPHI_BLOCK_DBG=$d timeout 300 python scripts/block_time.py 2>&1 | tail -1; done
