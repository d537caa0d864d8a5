#!/bin/bash
# round-2 GPU visit 6: full GPU tests, every workload's bench line, final ncu captures of the Ant step kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
grep -E "passed|failed|FAILED|fast-vs-exact" gpurun_out/pytest_gpu6.log | tail -12
run_bench() { n=$1; shift; timeout 400 python bench.py --steps 512 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r6b_$n.json 2> gpurun_out/r6b_$n.err; }
run_bench ant
run_bench anymal --workload anymal
run_bench humanoid --workload humanoid
run_bench cartpole --workload cartpole
run_bench hand --workload shadow_hand
for v in ant anymal humanoid cartpole hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r6b_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'rollout', d.get('rollout',{}).get('ms_per_step'))
except Exception as e:
    print('$v', 'failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 10 -c 1 -o gpurun_out/r2_ant_final -f python bench.py --steps 16 --warmup 3 --sets 2 --no-cpu-baseline --no-rollout > gpurun_out/ncu_final.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:quad_loco -s 60 -c 36 --csv --log-file gpurun_out/r2_ant_dram_rotating.csv python bench.py --steps 60 --warmup 3 --sets 18 --no-cpu-baseline --no-rollout > gpurun_out/ncu_dram.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
