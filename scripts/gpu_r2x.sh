#!/bin/bash
set -u
for i in 1 2; do
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), {k:(round(v['ms_per_evaluation'],4), round(v['value'],1)) for k,v in d['configs'].items()})"
done
