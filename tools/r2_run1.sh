#!/bin/bash
# round-2 GPU visit 1: microbench, GPU tests, Ant bench (quad 128 / quad 64 / generic), ncu of the quad kernel
mkdir -p gpurun_out
./tools/ubench > gpurun_out/ubench.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in q128 q64 generic; do
  case $v in
    q128) export B2G_QUAD_BLOCK=128; unset B2G_NO_QUAD;;
    q64) export B2G_QUAD_BLOCK=64; unset B2G_NO_QUAD;;
    generic) unset B2G_QUAD_BLOCK; export B2G_NO_QUAD=1;;
  esac
  timeout 300 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/r2_ant_$v.json 2> gpurun_out/r2_ant_$v.err
done
unset B2G_NO_QUAD; export B2G_QUAD_BLOCK=128
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_loco -s 40 -c 1 -o gpurun_out/r2_ant_quad_v1 -f python bench.py --steps 30 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_quad.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for v in q128 q64 generic; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_ant_$v.json').read().strip().splitlines()[-1])
    print('$v', 'us/step', d['ms_per_step']*1e3, 'b2b', d['back_to_back']['ms_per_step']*1e3, 'e2e', d['e2e']['ms_per_step']*1e3, 'frac', d['roofline']['frac'])
except Exception as e:
    print('$v', 'failed', e)
PY
done
