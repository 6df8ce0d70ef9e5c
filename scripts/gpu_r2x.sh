#!/bin/bash
set -u
O=gpurun_out/r2y; mkdir -p $O; rm -f $O/probe_small.txt
timeout 600 python -m pytest tests/test_gpu_persistent.py -x -q 2>&1 | tail -4
for B in "64 8" "128 8" "192 8" "256 8" "384 8" "8 4"; do
for nc in 2 4; do
echo "== B N = $B, T=100, chains $nc" | tee -a $O/probe_small.txt; PROBE_CHAINS=$nc timeout 300 python scripts/pipe_probe.py $B 100 10 2>&1 | tee -a $O/probe_small.txt | tail -3
done; done
