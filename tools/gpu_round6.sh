#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r6_tests.log 2>&1
SD_SAMPLER_STREAMS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_s1.log 2>&1
SD_SAMPLER_STREAMS=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_s2.log 2>&1
SD_SAMPLER_STREAMS=4 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_s4.log 2>&1
SD_SAMPLER_STREAMS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 40 --csv --log-file gpurun_out/r6_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r6_ncu_list.log 2>&1
for f in gpurun_out/r6_*.log; do echo "=== $f"; tail -n 5 $f | cut -c1-1700; done
