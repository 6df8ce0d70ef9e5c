#!/bin/bash
# multi-GPU visit: NCCL test of the sharded sampler + bench at N GPUs (weak headline + strong sub-dict)
set -u
N=${1:-2}
O=gpurun_out/r2g_n$N; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
echo "== multi-gpu test"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tee $O/pytest_multi.log | tail -5
echo "== bench N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>$O/bench.err | tee $O/bench.json | tail -2; tail -3 $O/bench.err
