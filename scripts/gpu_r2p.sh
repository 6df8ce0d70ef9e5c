#!/bin/bash
# persistent small-shard path: correctness (debug build traps instead of hanging), then timing
set -u
O=gpurun_out/r2p; mkdir -p $O
echo "== persistent tests"; timeout 300 python -m pytest tests/test_gpu_persistent.py -x -q 2>&1 | tee $O/pytest_persist.log | tail -15
echo "== timing (CCSP_PERSIST=0)"; CCSP_PERSIST=0 timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_off.jsonl | cut -c1-200
echo "== timing (persistent where eligible)"; timeout 300 python scripts/bench_configs.py 2>&1 | tee $O/configs_on.jsonl | cut -c1-200
