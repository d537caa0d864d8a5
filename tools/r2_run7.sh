#!/bin/bash
# round-2 GPU visit 7: arithmetic-form variants of the Ant kernel (A/B/A/B), self-collision extent measurement
mkdir -p gpurun_out
for rep in 1 2; do
for v in 000 001 101 111; do
  B2G_LIB=$PWD/variants/libb200gym_$v.so timeout 300 python bench.py --steps 1024 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/r7b_ant_$v.$rep.json 2> gpurun_out/r7b_ant_$v.$rep.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r7b_ant_$v.$rep.json').read().strip().splitlines()[-1])
    print('$v rep $rep', 'api us', round(d['ms_per_step']*1e3,3), 'dev us', round(d['device_only']['ms_per_step']*1e3,3), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,3))
except Exception as e:
    print('$v', 'failed', e)
PY
done
done
for v in 000 111; do
  B2G_LIB=$PWD/variants/libb200gym_$v.so timeout 300 python bench.py --workload anymal --steps 512 --warmup 5 --no-cpu-baseline > gpurun_out/r7b_anymal_$v.json 2> gpurun_out/r7b_anymal_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r7b_anymal_$v.json').read().strip().splitlines()[-1])
print('anymal $v', 'api us', round(d['ms_per_step']*1e3,3), 'dev us', round(d['device_only']['ms_per_step']*1e3,3), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,3))
PY
done
timeout 600 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision" > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
grep -E "passed|failed|overlapping|Error" gpurun_out/pytest_gpu7.log | tail
