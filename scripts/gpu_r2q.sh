#!/bin/bash
set -u
O=gpurun_out/r2q; mkdir -p $O
echo "== persistent tests"; timeout 300 python -m pytest tests/test_gpu_persistent.py -x -q 2>&1 | tee $O/pytest_persist.log | tail -5
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee $O/pytest_gpu.log | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -4
echo "== timing default"; timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_default.jsonl | cut -c1-200
echo "== timing forced persistent"; CCSP_PERSIST=1 timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_persist1.jsonl | cut -c1-200
echo "== timing persistent off"; CCSP_PERSIST=0 timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_persist0.jsonl | cut -c1-200
