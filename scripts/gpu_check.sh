#!/bin/bash
# Runs on the B200 box (via gpurun): GPU tests, smoke, a short bench, and an ncu launch list.
# Usage: scripts/gpu_check.sh [bench args...]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -25
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke 2>&1 | tee gpurun_out/smoke.log | tail -8
echo "== bench $*" ; timeout 1500 python bench.py "$@" 2>gpurun_out/bench.err | tee gpurun_out/bench.json | tail -3
tail -5 gpurun_out/bench.err
