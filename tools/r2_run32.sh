#!/bin/bash
# round-2 GPU visit 32: kin kernel with 16-lane groups for small articulations (tests, bench)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "kin or compat_gym_jac" > gpurun_out/pytest_gpu32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu32.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu32.log | tail -6
for mdl in "humanoid 8192" "ant 16384" "shadow_hand 4096" "anymal 4096"; do set -- $mdl; timeout 120 python tools/kin_bench.py --model $1 --envs $2 2>&1 | tail -1 | tee gpurun_out/kin_bench_$1.json | cut -c1-330; done
