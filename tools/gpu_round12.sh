#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r12_tests.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r12_bench.log 2>&1
SD_SAMPLER_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 5 -o gpurun_out/p2_conv_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r12_ncu.log 2>&1
SD_SAMPLER_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 848 -c 840 --csv --log-file gpurun_out/p2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r12_list.log 2>&1
for f in gpurun_out/r12_*.log; do echo "=== $f"; tail -n 6 $f | cut -c1-2200; done
