#!/bin/bash
# round-2 visit A: new checker tests first, then the whole GPU suite, smoke, bench
set -u
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
echo "== checker tests"; timeout 600 python -m pytest tests/test_gpu_checker.py -x -q 2>&1 | tee $O/pytest_checker.log | tail -15
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tee $O/pytest_gpu.log | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -5
echo "== bench"; timeout 900 python bench.py 2>$O/bench.err | tee $O/bench.json | tail -2; tail -3 $O/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$O/benchref.err | tee $O/benchref.json | tail -2; tail -3 $O/benchref.err
