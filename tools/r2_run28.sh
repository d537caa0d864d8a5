#!/bin/bash
# round-2 GPU visit 28: pen / egg after the object's angular-speed clamp; hand regression
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "hand or force or rounded or egg" > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu28.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_gpu28.log | tail -12
