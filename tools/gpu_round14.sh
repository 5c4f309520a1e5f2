#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r14_bench.log 2>&1
timeout 600 python bench.py --workload ref16 --steps 5 --warmup 3 --cpu-batch 8 > gpurun_out/r14_ref16.log 2>&1
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r14_cfg3.log 2>&1
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r14_cfg4.log 2>&1
for f in gpurun_out/r14_*.log; do echo "=== $f"; tail -n 3 $f | cut -c1-1500; done
