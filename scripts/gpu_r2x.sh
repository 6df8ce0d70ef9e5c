#!/bin/bash
set -u
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bf16x3" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs 2>$O/bench.err | tee $O/bench_orderB.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'clocks')}, d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['kernels']['avg_ms'])"
tail -2 $O/bench.err
