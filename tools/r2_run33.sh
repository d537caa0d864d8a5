#!/bin/bash
# round-2 GPU visit 33: the final library -- what the round-end driver runs (full GPU tests, smoke, default bench) + the ncu capture
# of the final Jacobian / mass-matrix kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu_final.log | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py 2>&1 | tail -1 > gpurun_out/final_ant.json; cut -c1-330 gpurun_out/final_ant.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kin_tensors -c 1 -o gpurun_out/kin_humanoid python tools/kin_bench.py --one > gpurun_out/ncu_kin.log 2>&1
python tools/ncu_summary.py gpurun_out/kin_humanoid.ncu-rep kin_tensors gpurun_out/r2_kin_humanoid_ncu_summary.json "Humanoid 8192 envs, J (16,6,27) + M (27,27), one launch; final kernel" 2>&1 | tail -1
python tools/ncu_lines.py gpurun_out/kin_humanoid.ncu-rep isaacgymenvs_b200/libb200gym.so kin_tensors 16 > gpurun_out/r2_kin_humanoid_lines.txt 2>&1; head -8 gpurun_out/r2_kin_humanoid_lines.txt
rm -f gpurun_out/kin_humanoid.ncu-rep
