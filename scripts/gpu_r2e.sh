#!/bin/bash
set -u
O=gpurun_out/r2e; mkdir -p $O
timeout 2400 python scripts/train_fixture.py --steps 40000 --eval-every 5000 --timesteps 1000 --out $O/ckpt 2>&1 | grep -v "^Step\|training completed" | tee $O/train_fixture.log | tail -30
ls -la $O/ckpt
