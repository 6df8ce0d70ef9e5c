#!/bin/bash
set -u
O=gpurun_out/r2f; mkdir -p $O
echo "== ebm tests"; timeout 900 python -m pytest tests/test_gpu_ebm.py -x -q 2>&1 | tee $O/pytest_ebm.log | tail -15
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee $O/pytest_gpu.log | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -5
echo "== bench"; timeout 900 python bench.py 2>$O/bench.err | tee $O/bench.json | tail -2; tail -3 $O/bench.err
