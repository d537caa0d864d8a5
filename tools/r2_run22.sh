#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
timeout 300 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "rollout_equals_k_single_steps and 16384" > gpurun_out/pytest_gpu22_$i.log 2>&1
grep -E "passed|failed|rollout != steps" gpurun_out/pytest_gpu22_$i.log | tail -2
done
