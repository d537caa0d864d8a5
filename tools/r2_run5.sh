#!/bin/bash
# round-2 GPU visit 5: the K-step rollout (tests, bench leg, ncu), fast-trig bound, anymal leg ordering experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -k "rollout or fast_trig" > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu5.log | tail -12
timeout 400 python bench.py --steps 512 --warmup 5 --no-cpu-baseline > gpurun_out/r5b_ant.json 2> gpurun_out/r5b_ant.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r5b_ant.json').read().strip().splitlines()[-1])
    print('ant api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'rollout us/step', round(d['rollout']['ms_per_step']*1e3,2))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r5b_ant.err').read()[-2000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_rollout -s 3 -c 1 -o gpurun_out/r2_ant_rollout_v1 -f python bench.py --steps 16 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_rollout1.log 2>&1
tail -3 gpurun_out/ncu_rollout1.log
