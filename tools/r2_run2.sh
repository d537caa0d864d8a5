#!/bin/bash
# round-2 GPU visit 2: all GPU tests, benches of every workload (+ Ant variants), ncu of the quad / anymal kernels, PPO
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
run_bench() { # name, extra args...
  n=$1; shift
  timeout 400 python bench.py --steps 500 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2b_$n.json 2> gpurun_out/r2b_$n.err
}
export B2G_QUAD_BLOCK=128
run_bench ant_q128
B2G_QUAD_NO_SPEC=1 run_bench ant_q128_nospec
B2G_QUAD_BLOCK=64 run_bench ant_q64
B2G_NO_QUAD=1 run_bench ant_generic
run_bench humanoid --workload humanoid
run_bench anymal --workload anymal
B2G_NO_QUAD=1 run_bench anymal_generic --workload anymal
run_bench cartpole --workload cartpole
run_bench hand --workload shadow_hand
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 40 -c 1 -o gpurun_out/r2_ant_quad_v2 -f python bench.py --steps 30 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_quad2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:anymal -s 12 -c 2 -o gpurun_out/r2_anymal_v2 -f python bench.py --workload anymal --steps 12 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_anymal2.log 2>&1
timeout 900 python tools/train_ppo.py --task Ant --num-envs 4096 --epochs 400 --out gpurun_out/r2_ppo_ant.json > gpurun_out/ppo_ant.log 2>&1
tail -15 gpurun_out/pytest_gpu2.log
for v in ant_q128 ant_q128_nospec ant_q64 ant_generic humanoid anymal anymal_generic cartpole hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2b_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'host_issue us', round(d['api']['host_issue_ms_per_step']*1e3,2))
except Exception as e:
    print('$v', 'failed', e)
PY
done
tail -3 gpurun_out/ppo_ant.log
