#!/bin/bash
# round-2 GPU visit 12: final state -- full GPU tests, all bench lines, ncu of the lean Ant kernel, Humanoid kernel with link-link contact
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu12.log
grep -E "passed|failed|FAILED|fast-vs-exact|overlap" gpurun_out/pytest_gpu12.log | tail -12
timeout 600 python bench.py --steps 1024 --warmup 5 > gpurun_out/r12_bench_ant.json 2> gpurun_out/r12_bench_ant.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r12_bench_reference.json 2> gpurun_out/r12_bench_reference.err
for w in anymal humanoid cartpole shadow_hand; do
  timeout 600 python bench.py --workload $w --steps 512 --warmup 5 > gpurun_out/r12_bench_$w.json 2> gpurun_out/r12_bench_$w.err
done
for v in ant anymal humanoid cartpole shadow_hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r12_bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'flushed', round(d['l2_flushed']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'cpu', round(d['cpu_baseline']['value']), 'rollout', d.get('rollout',{}).get('ms_per_step'))
except Exception as e:
    print('$v', 'failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 10 -c 1 -o gpurun_out/r2_ant_final -f python bench.py --steps 16 --warmup 3 --sets 2 --no-cpu-baseline --no-rollout > gpurun_out/ncu_final.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:quad_loco -s 60 -c 36 --csv --log-file gpurun_out/r2_ant_dram_rotating.csv python bench.py --steps 60 --warmup 3 --sets 18 --no-cpu-baseline --no-rollout > gpurun_out/ncu_dram.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ncu_launches_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_anymal -s 6 -c 1 -o gpurun_out/r2_anymal_final -f python bench.py --workload anymal --steps 12 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_anymal_final.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
