#!/bin/bash
set -u
O=gpurun_out/r2v; mkdir -p $O
echo "== parity (bf16 modes)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_persistent.py -q -x 2>&1 | tail -4
echo "== timing default"; timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_default.jsonl | cut -c1-200
echo "== timing persistent off"; CCSP_PERSIST=0 timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_persist0.jsonl | cut -c1-200
echo "== timing forced persistent"; CCSP_PERSIST=1 timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_persist1.jsonl | cut -c1-200
CCSP_PERSIST_TRACE=1 timeout 120 python scripts/persist_probe.py 8 4 | tail -3 | cut -c1-400
