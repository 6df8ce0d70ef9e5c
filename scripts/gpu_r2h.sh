#!/bin/bash
# CUDA-graph experiment: per-evaluation time of the sampling loop with and without graph replay, all configs (T = 40)
set -u
O=gpurun_out/r2h; mkdir -p $O

echo "== CUDA graph"; CCSP_GRAPH=1 timeout 600 python scripts/bench_configs.py 2>&1 | tee $O/configs_graph.jsonl
echo "== parity under graph"; CCSP_GRAPH=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "trajectory_vs_reference_golden and bf16x3" 2>&1 | tail -3
