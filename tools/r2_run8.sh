#!/bin/bash
# round-2 GPU visit 8: full GPU tests, every workload's bench line (with CPU baselines), reference arm, final ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
grep -E "passed|failed|FAILED|fast-vs-exact|overlapping" gpurun_out/pytest_gpu8.log | tail -12
timeout 600 python bench.py --steps 1024 --warmup 5 > gpurun_out/r8_bench_ant.json 2> gpurun_out/r8_bench_ant.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r8_bench_reference.json 2> gpurun_out/r8_bench_reference.err
for w in anymal humanoid cartpole shadow_hand; do
  timeout 600 python bench.py --workload $w --steps 512 --warmup 5 > gpurun_out/r8_bench_$w.json 2> gpurun_out/r8_bench_$w.err
done
for v in ant anymal humanoid cartpole shadow_hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r8_bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'flushed', round(d['l2_flushed']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'cpu', round(d['cpu_baseline']['value']), 'rollout', d.get('rollout',{}).get('ms_per_step'))
except Exception as e:
    print('$v', 'failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 10 -c 1 -o gpurun_out/r2_ant_final -f python bench.py --steps 16 --warmup 3 --sets 2 --no-cpu-baseline --no-rollout > gpurun_out/ncu_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_rollout -s 2 -c 1 -o gpurun_out/r2_ant_rollout -f python bench.py --steps 64 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_rollout.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:quad_loco -s 60 -c 36 --csv --log-file gpurun_out/r2_ant_dram_rotating.csv python bench.py --steps 60 --warmup 3 --sets 18 --no-cpu-baseline --no-rollout > gpurun_out/ncu_dram.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/train_ppo.py --task Ant --num-envs 4096 --epochs 400 --out gpurun_out/r2_ppo_ant.json > gpurun_out/ppo_ant.log 2>&1; tail -1 gpurun_out/ppo_ant.log
ls -la gpurun_out/*.ncu-rep | tail -3
