#!/bin/bash
# round-2 GPU visit 31: kin kernel with 128-bit packed scratch rows (tests, bench, ncu)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "kin or compat_gym_jac" > gpurun_out/pytest_gpu31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu31.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu31.log | tail -6
for mdl in "humanoid 8192" "ant 16384" "shadow_hand 4096"; do set -- $mdl; timeout 120 python tools/kin_bench.py --model $1 --envs $2 2>&1 | tail -1 | tee gpurun_out/kin_bench_$1.json | cut -c1-330; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kin_tensors -c 1 -o gpurun_out/kin_humanoid python tools/kin_bench.py --one > gpurun_out/ncu_kin.log 2>&1
python tools/ncu_summary.py gpurun_out/kin_humanoid.ncu-rep kin_tensors gpurun_out/r2_kin_humanoid_ncu_summary.json "Humanoid 8192 envs, J (16,6,27) + M (27,27), one launch; 128-bit packed scratch" 2>&1 | tail -1
python tools/ncu_lines.py gpurun_out/kin_humanoid.ncu-rep isaacgymenvs_b200/libb200gym.so kin_tensors 14 > gpurun_out/r2_kin_humanoid_lines.txt 2>&1; head -16 gpurun_out/r2_kin_humanoid_lines.txt
rm -f gpurun_out/kin_humanoid.ncu-rep
