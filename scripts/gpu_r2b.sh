#!/bin/bash
# round-2 visit B: training-step tests, checker tests, then a short fixture training run with solved rates
set -u
O=gpurun_out/r2b; mkdir -p $O
echo "== train tests"; timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tee $O/pytest_train.log | tail -25
echo "== checker tests"; timeout 600 python -m pytest tests/test_gpu_checker.py -x -q 2>&1 | tee $O/pytest_checker.log | tail -8
echo "== fixture training"; timeout 1200 python scripts/train_fixture.py --steps 2000 --eval-every 500 --timesteps 200 --out $O/ckpt 2>&1 | tee $O/train_fixture.log | tail -20
rm -f $O/ckpt/model-fixture.pt
ls -la $O/ckpt
