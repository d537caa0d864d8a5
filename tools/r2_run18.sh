#!/bin/bash
# round-2 GPU visit 18: final verification of the committed state + Humanoid learning curves (with / without link-link contact)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu18.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu18.log | tail -8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r18_bench_default.json 2> gpurun_out/r18_bench_default.err; tail -c 400 gpurun_out/r18_bench_default.json
timeout 400 python tools/train_ppo.py --task Humanoid --num-envs 4096 --epochs 600 --horizon 32 --units 400,200,100 --lr 5e-4 --mini-epochs 5 --critic-coef 4 --out gpurun_out/r2_ppo_humanoid.json > gpurun_out/ppo_humanoid.log 2>&1; tail -1 gpurun_out/ppo_humanoid.log | cut -c1-420
timeout 500 python tools/train_ppo.py --task Humanoid --num-envs 4096 --epochs 600 --horizon 32 --units 400,200,100 --lr 5e-4 --mini-epochs 5 --critic-coef 4 --self-collision --out gpurun_out/r2_ppo_humanoid_selfcollision.json > gpurun_out/ppo_humanoid_sc.log 2>&1; tail -1 gpurun_out/ppo_humanoid_sc.log | cut -c1-420
