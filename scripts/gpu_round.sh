#!/bin/bash
# One GPU-box visit: kernel harness (numerics + ablations), bench, ncu launch list, GPU tests, smoke.
# Usage: scripts/gpu_round.sh [tag]
set -u
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
echo "== harness"; timeout 300 diffusion_ccsp_b200/lib/tc_gemm_test perf > $O/harness.txt 2>&1; tail -50 $O/harness.txt
echo "== bench"; timeout 600 python bench.py 2>$O/bench.err | tee $O/bench.json | tail -2; tail -3 $O/bench.err
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --timesteps 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_bench.log 2>&1; tail -2 $O/ncu_bench.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tee $O/pytest_gpu.log | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -5
