#!/bin/bash
# round-2 GPU visit 3: all GPU tests, Ant / AnymalTerrain benches, ncu of the quad kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
run_bench() { n=$1; shift; timeout 400 python bench.py --steps 500 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r3b_$n.json 2> gpurun_out/r3b_$n.err; }
export B2G_QUAD_BLOCK=128
run_bench ant_q128
B2G_QUAD_BLOCK=64 run_bench ant_q64
run_bench anymal --workload anymal
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 40 -c 1 -o gpurun_out/r2_ant_quad_v3 -f python bench.py --steps 30 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_quad3.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu3.log | tail -12
for v in ant_q128 ant_q64 anymal; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3b_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'host_issue us', round(d['api']['host_issue_ms_per_step']*1e3,2))
except Exception as e:
    print('$v', 'failed', e)
PY
done
