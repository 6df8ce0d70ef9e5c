#!/bin/bash
# Re-validation after a container rebuild: full GPU suite, smoke, default bench, reference arm.
set -u
O=gpurun_out/${1:-r2w}; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tee $O/pytest_gpu.log | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -4
echo "== bench"; timeout 900 python bench.py 2>$O/bench.err | tee $O/bench.json | cut -c1-600; tail -3 $O/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>$O/ref.err | tee $O/bench_ref.json | cut -c1-600
