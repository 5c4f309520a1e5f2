#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r10_tests.log 2>&1
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r10_cfg4.log 2>&1
SD_TC_TCHUNK=0 timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r10_cfg4_nochunk.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench.log 2>&1
for f in gpurun_out/r10_*.log; do echo "=== $f"; tail -n 12 $f | cut -c1-1800; done
