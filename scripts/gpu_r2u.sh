#!/bin/bash
set -u
O=gpurun_out/r2u; mkdir -p $O
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checker.py -q -k "benchmarked_workload or order_do_not_matter or full_size" 2>&1 | tee $O/pytest_new.log | tail -6
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tee $O/pytest_gpu.log | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $O/smoke.log | tail -3
echo "== bench"; timeout 900 python bench.py 2>$O/bench.err | tee $O/bench.json | cut -c1-300; tail -2 $O/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$O/benchref.err | tee $O/benchref.json | cut -c1-200
