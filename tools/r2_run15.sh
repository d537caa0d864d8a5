#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loco_step -s 35 -c 1 -o /tmp/hum_self -f python tools/humanoid_self_probe.py > gpurun_out/ncu_hum_self.log 2>&1
python tools/ncu_summary.py /tmp/hum_self.ncu-rep loco_step gpurun_out/r2_humanoid_self_ncu_summary.json "Humanoid 8192 envs with link-link contact (env.selfCollision=True)" > /dev/null 2>&1
python tools/ncu_lines.py /tmp/hum_self.ncu-rep isaacgymenvs_b200/libb200gym.so loco_step_kernelILi4ELb0ELb1ELi64ELb1ELb0ELb1 60 > gpurun_out/r2_humanoid_self_lines.txt 2>&1
head -40 gpurun_out/r2_humanoid_self_lines.txt
grep -E "duration|inst_executed.sum|registers|warps_active.avg.per" gpurun_out/r2_humanoid_self_ncu_summary.json
