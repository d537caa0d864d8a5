#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/rollout_probe.py 2>&1 | grep -v Warn | tail -24
