#!/bin/bash
# compute-sanitizer memcheck over the small probe (all kernel families incl. the sampling kernels)
set -u
O=gpurun_out/sanitize; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 $O/memcheck.log | cut -c1-200
