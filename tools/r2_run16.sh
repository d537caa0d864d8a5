#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision or interpenetrate" > gpurun_out/pytest_gpu16a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu16a.log
grep -E "passed|failed|FAILED|overlap|^E  " gpurun_out/pytest_gpu16a.log | tail -10
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loco_step -s 35 -c 1 -o /tmp/hum_self -f python tools/humanoid_self_probe.py > gpurun_out/ncu_hum_self.log 2>&1
python tools/ncu_summary.py /tmp/hum_self.ncu-rep loco_step gpurun_out/r2_humanoid_self_ncu_summary.json "Humanoid 8192 envs with link-link contact (env.selfCollision=True)" > /dev/null 2>&1
python tools/ncu_lines.py /tmp/hum_self.ncu-rep isaacgymenvs_b200/libb200gym.so loco_step_kernelILi4ELb0ELb1ELi64ELb1ELb0ELb1 40 > gpurun_out/r2_humanoid_self_lines.txt 2>&1
head -24 gpurun_out/r2_humanoid_self_lines.txt
grep -E "duration|inst_executed.sum|issue_active" gpurun_out/r2_humanoid_self_ncu_summary.json
