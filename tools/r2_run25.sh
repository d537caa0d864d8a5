#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu25.log
tail -4 gpurun_out/pytest_gpu25.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py 2>&1 | tail -1 > gpurun_out/bench25_ant.json; cut -c1-400 gpurun_out/bench25_ant.json
