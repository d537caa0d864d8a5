#!/bin/bash
# round-2 GPU visit 19: link-link contact with the stable (reduced-mass) gains: tests, cost, Humanoid learning with it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision or interpenetrate" > gpurun_out/pytest_gpu19a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu19a.log
grep -E "passed|failed|FAILED|overlap|^E  " gpurun_out/pytest_gpu19a.log | tail -10
cat > /tmp/hum_sc.py <<'PY'
import sys, os, time, torch
sys.path.insert(0, '.')
import isaacgymenvs_b200
from isaacgymenvs_b200 import config
for on in (False, True):
    cfg = config.builtin_cfg("Humanoid", {"sim_device": "cuda:0", "rl_device": "cuda:0"}); cfg["task"]["env"]["selfCollision"] = on
    env = isaacgymenvs_b200.make(seed=1, task="Humanoid", num_envs=8192, sim_device="cuda:0", rl_device="cuda:0", headless=True, cfg=cfg)
    g = torch.Generator(device="cuda:0"); g.manual_seed(1)
    acts = [torch.rand((8192, env.num_acts), device="cuda:0", generator=g) * 2 - 1 for _ in range(16)]
    for k in range(20): env.step(acts[k % 16])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(300): env.step(acts[k % 16])
    t1.record(); torch.cuda.synchronize()
    print("humanoid 8192 envs selfCollision", on, "us/step", round(t0.elapsed_time(t1) / 300 * 1e3, 2), flush=True)
PY
timeout 300 python /tmp/hum_sc.py 2>&1 | grep -v Warning | tail -2
timeout 500 python tools/train_ppo.py --task Humanoid --num-envs 4096 --epochs 600 --horizon 32 --units 400,200,100 --lr 5e-4 --mini-epochs 5 --critic-coef 4 --self-collision --out gpurun_out/r2_ppo_humanoid_selfcollision.json > gpurun_out/ppo_humanoid_sc.log 2>&1; tail -1 gpurun_out/ppo_humanoid_sc.log | cut -c1-420
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu19.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu19.log | tail -6
