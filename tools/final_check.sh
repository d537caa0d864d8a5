#!/bin/bash
# Run from the repo root (./tools/final_check.sh).  One GPU-box pass over everything the round-end driver runs, plus the evidence files copied into profiles/.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in ant humanoid anymal cartpole shadow_hand; do
  timeout 200 python bench.py --workload $w --steps 1000 --warmup 5 2>&1 | tail -1 > gpurun_out/final_$w.json
  python -c "
import json; d=json.load(open('gpurun_out/final_$w.json'))
print('$w', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['ms_per_step']*1e3,1), 'us b2b', round(d['back_to_back']['ms_per_step']*1e3,1), 'flushed', round(d['l2_flushed']['ms_per_step']*1e3,1), 'frac', round(d['roofline']['frac'],4), 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks'])"
done
timeout 200 python bench.py --impl reference --steps 30 --warmup 2 2>&1 | tail -1 > gpurun_out/final_reference.json; cut -c1-300 gpurun_out/final_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/r1_final_launches.csv python bench.py --steps 40 --warmup 5 --sets 2 --no-cpu-baseline > /dev/null 2>&1
tail -5 gpurun_out/r1_final_launches.csv | cut -c1-200
