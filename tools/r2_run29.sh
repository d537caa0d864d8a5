#!/bin/bash
# round-2 GPU visit 29: what the round-end driver runs (full GPU tests, smoke, bench) + bench lines of every workload, the
# reference arm, and the new ShadowHand options under a learner
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu_final.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in ant humanoid anymal cartpole shadow_hand; do
  timeout 200 python bench.py --workload $w --steps 600 --warmup 5 2>&1 | tail -1 > gpurun_out/final_$w.json
  python -c "
import json; d=json.load(open('gpurun_out/final_$w.json'))
print('$w', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), d['clocks'])"
done
timeout 200 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 > gpurun_out/final_reference.json; cut -c1-200 gpurun_out/final_reference.json
timeout 150 python tools/train_ppo.py --task ShadowHand --num-envs 8192 --epochs 150 --horizon 8 --units 512,512,256,128 --lr 5e-4 --mini-epochs 5 --critic-coef 4 --env objectType=pen,forceScale=1.0 --out gpurun_out/r2_ppo_shadow_hand_pen_forces.json > gpurun_out/ppo_pen.log 2>&1; tail -1 gpurun_out/ppo_pen.log | cut -c1-300
timeout 150 python tools/train_ppo.py --task ShadowHand --num-envs 8192 --epochs 150 --horizon 8 --units 512,512,256,128 --lr 5e-4 --mini-epochs 5 --critic-coef 4 --env objectType=egg --out gpurun_out/r2_ppo_shadow_hand_egg.json > gpurun_out/ppo_egg.log 2>&1; tail -1 gpurun_out/ppo_egg.log | cut -c1-300
