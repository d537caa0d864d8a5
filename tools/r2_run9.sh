#!/bin/bash
# round-2 GPU visit 9 (8 GPUs of one box): weak and strong scaling of the Ant step, ShadowHand 32768 envs over 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 512 --warmup 5 --no-cpu-baseline > gpurun_out/r9_ant_n1.json 2> gpurun_out/r9_ant_n1.err
timeout 400 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 512 --warmup 5 --no-cpu-baseline > gpurun_out/r9_ant_n8.json 2> gpurun_out/r9_ant_n8.err
timeout 400 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 512 --warmup 5 --no-cpu-baseline --scaling strong > gpurun_out/r9_ant_n8_strong.json 2> gpurun_out/r9_ant_n8_strong.err
timeout 400 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 256 --warmup 5 --no-cpu-baseline --workload shadow_hand > gpurun_out/r9_hand_n8.json 2> gpurun_out/r9_hand_n8.err
timeout 400 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 512 --warmup 5 --no-cpu-baseline > gpurun_out/r9_ant_n2.json 2> gpurun_out/r9_ant_n2.err
for v in ant_n1 ant_n2 ant_n8 ant_n8_strong hand_n8; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r9_$v.json').read().strip().splitlines()[-1])
    print('$v', 'value', round(d['value']/1e9,4), 'G env-steps/s', 'us/step', round(d['ms_per_step']*1e3,2), 'n_gpus', d['n_gpus'], d['scaling'], d['config']['num_envs_total'])
except Exception as e:
    print('$v', 'failed', e); print(open('gpurun_out/r9_$v.err').read()[-1500:])
PY
done
