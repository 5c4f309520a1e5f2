#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r5_tests.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 60 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r5_ncu_list.log 2>&1
for f in gpurun_out/r5_*.log; do echo "=== $f"; tail -n 6 $f | cut -c1-1800; done
