#!/bin/bash
# round-2 GPU visit 4: tests, every workload's bench, Ant CTA-size sweep, DRAM traffic in the rotating-set mode, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
run_bench() { n=$1; shift; timeout 400 python bench.py --steps 500 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r4b_$n.json 2> gpurun_out/r4b_$n.err; }
B2G_QUAD_BLOCK=32 run_bench ant_q32
B2G_QUAD_BLOCK=64 run_bench ant_q64
B2G_QUAD_BLOCK=128 run_bench ant_q128
run_bench anymal --workload anymal
run_bench humanoid --workload humanoid
run_bench cartpole --workload cartpole
run_bench hand --workload shadow_hand
# the driver's own invocation shape (default workload, cpu baseline leg included) + the reference arm
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r4_bench_reference.json 2> gpurun_out/r4_bench_reference.err
# DRAM traffic of the step kernel in the rotating-set (HBM-cold data) mode: no cache flush by ncu, 18 sets
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:quad_loco -s 60 -c 36 --csv --log-file gpurun_out/r2_ant_dram_rotating.csv python bench.py --steps 60 --warmup 3 --sets 18 --no-cpu-baseline > gpurun_out/ncu_dram.log 2>&1
# launch list of the default bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:anymal -s 12 -c 2 -o gpurun_out/r2_anymal_v3 -f python bench.py --workload anymal --steps 12 --warmup 3 --sets 2 --no-cpu-baseline > gpurun_out/ncu_anymal3.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu4.log | tail -12
for v in ant_q32 ant_q64 ant_q128 anymal humanoid cartpole hand; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r4b_$v.json').read().strip().splitlines()[-1])
    print('$v', 'api us', round(d['ms_per_step']*1e3,2), 'dev us', round(d['device_only']['ms_per_step']*1e3,2), 'b2b', round(d['back_to_back']['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'host_issue us', round(d['api']['host_issue_ms_per_step']*1e3,2))
except Exception as e:
    print('$v', 'failed', e)
PY
done
tail -c 600 gpurun_out/r4_bench_default.json; echo; tail -c 400 gpurun_out/r4_bench_reference.json
