#!/bin/bash
set -u
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_persistent.py -x -q 2>&1 | tail -5
for nc in 2 3; do
echo "== config 2, T=150, chains $nc"; PROBE_CHAINS=$nc timeout 300 python scripts/pipe_probe.py 1024 8 150 10 20 24 2>&1 | tee $O/probe_cfg2_c$nc.txt | tail -4
echo "== config 2, T=150, chains $nc, drain"; CCSP_PIPE_DRAIN=1 PROBE_CHAINS=$nc timeout 300 python scripts/pipe_probe.py 1024 8 150 10 24 2>&1 | tail -1
done
