#!/bin/bash
# same box: release build vs CCSP_DEBUG (bounded spin) build of the library, headline bench without the extras
set -u
O=gpurun_out/r2s; mkdir -p $O
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs"
for v in release debug release debug; do
  if [ $v = debug ]; then CCSP_DEBUG=1 python -m diffusion_ccsp_b200.build --force > /dev/null 2>&1; else python -m diffusion_ccsp_b200.build --force > /dev/null 2>&1; fi
  timeout 300 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'],1), 'scenes/s  edge', round(d['kernels']['avg_ms']['edge_l1'],4), 'node', round(d['kernels']['avg_ms']['node'],4), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w_max'))"
done
python -m diffusion_ccsp_b200.build --force > /dev/null 2>&1
