#!/bin/bash
# ncu on the GPU box: launch list + one --set full capture of the edge and node kernels.  Usage: scripts/gpu_prof.sh [tag]
set -u
TAG=${1:-prof}
O=gpurun_out/$TAG
mkdir -p $O
B="python bench.py --timesteps 6 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file $O/launches.csv $B > $O/ncu_list.log 2>&1; tail -1 $O/ncu_list.log | cut -c1-200
echo "== full capture"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge_fused2|k_node_tc' -s 20 -c 4 -o $O/prof $B > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log | cut -c1-200
ls -la $O
