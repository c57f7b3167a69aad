#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -k "wide or other_activation or cfg4 or cfg1 or lift or block" -s > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r2h_pytest.log | tail -25
