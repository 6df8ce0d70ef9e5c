#!/bin/bash
set -u
O=gpurun_out/r2t; mkdir -p $O
timeout 2400 python scripts/train_fixture.py --steps 60000 --batch 128 --max-tiles 5 --eval-every 10000 --timesteps 1000 --out $O/ckpt 2>&1 | grep -v "^Step\|training completed" | tee $O/train_fixture_n5.log | tail -30
