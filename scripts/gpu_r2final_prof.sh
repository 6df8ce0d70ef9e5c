#!/bin/bash
# final round-2 ncu visit: launch list of the bench command + one --set full capture of the sampling kernels
set -u
O=gpurun_out/r2final; mkdir -p $O
B="python bench.py --timesteps 6 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
echo "== launch list (bench)"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 140 --csv --log-file $O/launches_r2_final.csv $B > $O/ncu_list.log 2>&1; tail -1 $O/ncu_list.log | cut -c1-200
echo "== full capture (sampling kernels)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge_fused2|k_node_tc' -s 20 -c 4 -o $O/prof_sample $B > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log | cut -c1-200
ls -la $O
