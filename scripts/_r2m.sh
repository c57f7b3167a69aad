#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "transfer or golden or rational_layer or rollout or tensor_core_layer" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2m_pytest.log
python scripts/profile_bwd_kernel.py > gpurun_out/r2m_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:block_tail_bwd -s 2 -c 1 -f -o gpurun_out/r2m_bwd_za python scripts/profile_bwd_kernel.py > gpurun_out/r2m_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2m_ncu.log; ls -la gpurun_out/*.ncu-rep
