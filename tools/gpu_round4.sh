#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 420 --csv --log-file gpurun_out/r4_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_ncu_list.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r4_bench.log 2>&1
for f in gpurun_out/r4_*.log; do echo "=== $f"; tail -n 4 $f | cut -c1-2500; done
