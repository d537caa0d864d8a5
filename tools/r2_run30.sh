#!/bin/bash
# round-2 GPU visit 30: hand kernels after moving the object force out of registers (tests + bench)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "hand or force or rounded or egg" > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu30.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu30.log | tail -6
timeout 200 python bench.py --workload shadow_hand --steps 1000 --warmup 5 2>&1 | tail -1 > gpurun_out/final_shadow_hand.json
python -c "
import json; d=json.load(open('gpurun_out/final_shadow_hand.json'))
print('shadow_hand', round(d['ms_per_step']*1e3,1), 'us e2e', round(d['e2e']['ms_per_step']*1e3,1), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,1), 'dev', round(d['device_only']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4))"
