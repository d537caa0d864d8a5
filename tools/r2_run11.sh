#!/bin/bash
# round-2 GPU visit 11: lean Ant kernel, link-link contact with the bounding-sphere broad phase (Humanoid with / without)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision or interpenetrate" > gpurun_out/pytest_gpu11a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu11a.log
grep -E "passed|failed|FAILED|overlap|^E  " gpurun_out/pytest_gpu11a.log | tail -12
for v in 0 1; do
B2G_NO_LEAN=$v timeout 300 python bench.py --steps 1024 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/r11_ant_nolean$v.json 2> gpurun_out/r11_ant_nolean$v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r11_ant_nolean$v.json').read().strip().splitlines()[-1])
    print('ant B2G_NO_LEAN=$v', 'api us', round(d['ms_per_step']*1e3,3), 'dev us', round(d['device_only']['ms_per_step']*1e3,3), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,3))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r11_ant_nolean$v.err').read()[-1500:])
PY
done
cat > /tmp/hum_sc.py <<'PY'
import sys, time, torch
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
timeout 300 python /tmp/hum_sc.py 2>&1 | grep -v Warning | tail -3
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu11.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu11.log | tail -8
