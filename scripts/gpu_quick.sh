#!/bin/bash
# Short GPU-box visit: GPU tests, smoke, bench (no harness, no ncu).  Usage: scripts/gpu_quick.sh [tag] [bench args...]
set -u
TAG=${1:-quick}; shift || true
O=gpurun_out/$TAG
mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tee $O/pytest_gpu.log | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -5
echo "== bench"; timeout 600 python bench.py "$@" 2>$O/bench.err | tee $O/bench.json | tail -2; tail -3 $O/bench.err
