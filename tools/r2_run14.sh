#!/bin/bash
# round-2 GPU visit 14: link-link contact in idle slot cells (template-gated), Humanoid default path restored?
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision or interpenetrate or humanoid" > gpurun_out/pytest_gpu14a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu14a.log
grep -E "passed|failed|FAILED|overlap|^E  " gpurun_out/pytest_gpu14a.log | tail -14
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
    print("humanoid 8192 envs selfCollision", on, os.environ.get("B2G_SELF_APPENDED", ""), "us/step", round(t0.elapsed_time(t1) / 300 * 1e3, 2), flush=True)
PY
timeout 300 python /tmp/hum_sc.py 2>&1 | grep -v Warning | tail -2
B2G_SELF_APPENDED=1 timeout 300 python /tmp/hum_sc.py 2>&1 | grep -v Warning | tail -1
timeout 600 python bench.py --workload humanoid --steps 512 --warmup 5 > gpurun_out/r14_bench_humanoid.json 2> gpurun_out/r14_bench_humanoid.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r14_bench_humanoid.json').read().strip().splitlines()[-1])
print('humanoid bench', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2))
PY
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu14.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu14.log | tail -8
