#!/bin/bash
# round-2 GPU visit 13: cooperative link-link contact; final bench lines; ncu captures summarised ON THE BOX (reports are > 30 MB each)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "self_collision or interpenetrate" > gpurun_out/pytest_gpu13a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu13a.log
grep -E "passed|failed|FAILED|overlap|^E  " gpurun_out/pytest_gpu13a.log | tail -14
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
    print("humanoid 8192 envs selfCollision", on, "us/step", round(t0.elapsed_time(t1) / 300 * 1e3, 2), "block", env.sim.block_size() if hasattr(env.sim, "block_size") else "?", flush=True)
PY
timeout 300 python /tmp/hum_sc.py 2>&1 | grep -v Warning | tail -3
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu13.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu13.log | tail -8
timeout 600 python bench.py --steps 1024 --warmup 5 > gpurun_out/r13_bench_ant.json 2> gpurun_out/r13_bench_ant.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r13_bench_reference.json 2> gpurun_out/r13_bench_reference.err
for w in anymal humanoid cartpole shadow_hand; do
  timeout 600 python bench.py --workload $w --steps 512 --warmup 5 > gpurun_out/r13_bench_$w.json 2> gpurun_out/r13_bench_$w.err
done
for v in ant anymal humanoid cartpole shadow_hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r13_bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'flushed', round(d['l2_flushed']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'cpu', round(d['cpu_baseline']['value']), 'rollout', d.get('rollout',{}).get('ms_per_step'))
except Exception as e:
    print('$v', 'failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 10 -c 1 -o /tmp/r2_ant_final -f python bench.py --steps 16 --warmup 3 --sets 2 --no-cpu-baseline --no-rollout > gpurun_out/ncu_final.log 2>&1
python tools/ncu_summary.py /tmp/r2_ant_final.ncu-rep quad_loco gpurun_out/r2_ant_final_ncu_summary.json "Ant 16384 envs, quad_loco_kernel<2,3,64,false,LEAN=true> (ncu --set full --clock-control none, one launch, cold cache, serialised; bench.py --sets 2)" > /dev/null 2>&1
python tools/ncu_lines.py /tmp/r2_ant_final.ncu-rep isaacgymenvs_b200/libb200gym.so quad_loco_kernelILi2ELi3ELi64ELb0ELb1 70 > gpurun_out/r2_ant_final_lines.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_anymal -s 6 -c 1 -o /tmp/r2_anymal_final -f python bench.py --workload anymal --steps 12 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_anymal_final.log 2>&1
python tools/ncu_summary.py /tmp/r2_anymal_final.ncu-rep quad_anymal gpurun_out/r2_anymal_physics_ncu_summary.json "AnymalTerrain 4096 envs: quad_anymal_physics_kernel<true,128,DR=false>" > /dev/null 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:quad_loco -s 60 -c 36 --csv --log-file gpurun_out/r2_ant_dram_rotating.csv python bench.py --steps 60 --warmup 3 --sets 18 --no-cpu-baseline --no-rollout > gpurun_out/ncu_dram.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ncu_launches_full.log 2>&1
du -sh gpurun_out
