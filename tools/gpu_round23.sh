#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r23_tests.log 2>&1
timeout 600 python bench.py --workload ref16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r23_ref16.log 2>&1
for f in gpurun_out/r23_*.log; do echo "=== $f"; tail -n 3 $f | cut -c1-330; done
