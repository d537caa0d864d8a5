#!/bin/bash
# round-2 GPU visit 10: link-link contact (Humanoid) -- tests, Humanoid bench with and without it, full launch list of the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision or interpenetrate" > gpurun_out/pytest_gpu10a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10a.log
grep -E "passed|failed|FAILED|overlap|Error|assert" gpurun_out/pytest_gpu10a.log | tail -12
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu10.log | tail -12
timeout 400 python bench.py --workload humanoid --steps 512 --warmup 5 > gpurun_out/r10_bench_humanoid.json 2> gpurun_out/r10_bench_humanoid.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r10_bench_humanoid.json').read().strip().splitlines()[-1])
    print('humanoid (self-collision on)', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'cpu', round(d['cpu_baseline']['value']))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r10_bench_humanoid.err').read()[-1500:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ncu_launches_full.log 2>&1
tail -2 gpurun_out/ncu_launches_full.log | cut -c1-300
