#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bf16x3 or full_batch or full_size or benchmarked" 2>&1 | tail -3
for i in 1; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), d['kernels']['avg_ms'], {k:(round(v['ms_per_evaluation'],4), round(v['value'],1)) for k,v in d['configs'].items()})"
done
