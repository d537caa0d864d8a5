#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -k "rollout" > gpurun_out/pytest_gpu24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu24.log
grep -E "passed|failed|FAILED|rollout != steps" gpurun_out/pytest_gpu24.log | tail -10
