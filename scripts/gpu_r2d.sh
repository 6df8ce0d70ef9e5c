#!/bin/bash
set -u
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python scripts/train_fixture.py --steps 20000 --eval-every 5000 --timesteps 1000 --out $O/ckpt 2>&1 | grep -v "^Step\|training completed" | tee $O/train_fixture.log | tail -20
rm -f $O/ckpt/model-fixture.pt
timeout 900 python tests/tools/diagnose_sampling.py --ckpt $O/ckpt/denoise_fn_fp16.pt --scenes 4 --reference 2>&1 | tee $O/diagnose.log | tail -12
