#!/bin/bash
# round-2 GPU visit 34: does running the host-I/O Ant step in several waves overlap compute with the PCIe drain?  (e2e leg of bench.py)
mkdir -p gpurun_out
for c in 0 4 3 2; do
  B2G_HOSTIO_CTAS=$c timeout 100 python bench.py --steps 300 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/hostio_$c.json
  python -c "
import json; d=json.load(open('gpurun_out/hostio_$c.json'))
print('ctas/SM $c: e2e', round(d['e2e']['ms_per_step']*1e3,1), 'us  step', round(d['ms_per_step']*1e3,2), 'us')"
done
B2G_HOSTIO_CTAS=3 timeout 100 python -m pytest tests -m gpu -q -k "host_buffer" 2>&1 | tail -1
