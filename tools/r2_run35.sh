#!/bin/bash
# round-2 GPU visit 35: the final library once more through the full GPU suite and smoke
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu_final.log | tail -4
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
