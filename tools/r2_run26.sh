#!/bin/bash
# round-2 GPU visit 26: kinematic tensors (shadow_hand case, bench, ncu), random object forces (golden + physics), hand regression
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "kin or compat_gym_jac or hand or force" > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu26.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu26.log | tail -12
for mdl in "humanoid 8192" "ant 16384" "shadow_hand 4096"; do set -- $mdl; timeout 120 python tools/kin_bench.py --model $1 --envs $2 2>&1 | tail -1 | tee gpurun_out/kin_bench_$1.json | cut -c1-400; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kin_tensors -c 1 -o gpurun_out/kin_humanoid python tools/kin_bench.py --one > gpurun_out/ncu_kin.log 2>&1
python tools/ncu_summary.py gpurun_out/kin_humanoid.ncu-rep kin_tensors gpurun_out/r2_kin_humanoid_ncu_summary.json "Humanoid 8192 envs, J (16,6,27) + M (27,27), one launch" 2>&1 | tail -2
python tools/ncu_lines.py gpurun_out/kin_humanoid.ncu-rep isaacgymenvs_b200/libb200gym.so kin_tensors 30 > gpurun_out/r2_kin_humanoid_lines.txt 2>&1; head -30 gpurun_out/r2_kin_humanoid_lines.txt
rm -f gpurun_out/kin_humanoid.ncu-rep
