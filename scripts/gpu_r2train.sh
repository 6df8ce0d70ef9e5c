#!/bin/bash
set -u
O=gpurun_out/r2train; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_persistent.py -x -q 2>&1 | tail -2
echo "== A: 6..8 tiles, lr 2e-4"; timeout 600 python scripts/train_fixture.py --steps 30000 --eval-every 5000 --min-tiles 6 --lr 2e-4 --out $O/a 2>&1 | grep fixture | tee $O/a.log | cut -c1-200
echo "== B: all, lr 2e-4"; timeout 600 python scripts/train_fixture.py --steps 30000 --eval-every 5000 --lr 2e-4 --out $O/b 2>&1 | grep fixture | tee $O/b.log | cut -c1-200
