#!/bin/bash
set -u
O=gpurun_out/r2e; mkdir -p $O
echo "== train tests (kernel changes)"; timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -3
timeout 300 python scripts/train_probe.py 128 30 2>&1 | tail -1
timeout 2400 python scripts/train_fixture.py --steps 100000 --eval-every 10000 --timesteps 1000 --out $O/ckpt 2>&1 | grep -v "^Step\|training completed" | tee $O/train_fixture.log | tail -20
rm -f $O/ckpt/model-fixture.pt $O/ckpt/denoise_fn_fp16.pt
ls -la $O/ckpt
