#!/bin/bash
# round-2 GPU visit 21: per-env physical parameters in the generic sub-step; full tests; bench lines of the generic-path workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu21.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu21.log | tail -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in humanoid shadow_hand cartpole anymal; do
  timeout 600 python bench.py --workload $w --steps 512 --warmup 5 > gpurun_out/r21_bench_$w.json 2> gpurun_out/r21_bench_$w.err
done
timeout 600 python bench.py --steps 1024 --warmup 5 > gpurun_out/r21_bench_ant.json 2> gpurun_out/r21_bench_ant.err
for v in ant anymal humanoid cartpole shadow_hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r21_bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'flushed', round(d['l2_flushed']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'cpu', round(d['cpu_baseline']['value']))
except Exception as e:
    print('$v', 'failed', e)
PY
done
