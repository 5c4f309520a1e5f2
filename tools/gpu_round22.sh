#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r22_tests.log 2>&1
timeout 600 python bench.py --workload ref16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r22_ref16.log 2>&1
SD_TC_SMALL_BATCH_SPLIT=0 timeout 600 python bench.py --workload ref16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r22_ref16_nosplit.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r22_bench.log 2>&1
for f in gpurun_out/r22_*.log; do echo "=== $f"; tail -n 3 $f | cut -c1-330; done
