#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r9_tests.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench.log 2>&1
SD_SAMPLER_STREAMS=4 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench_s4.log 2>&1
SD_SAMPLER_GRAPH=0 SD_SAMPLER_STREAMS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 830 -c 14 --csv --log-file gpurun_out/r9_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r9_list.log 2>&1
for f in gpurun_out/r9_*.log; do echo "=== $f"; tail -n 12 $f | cut -c1-700; done
grep -v "^==" gpurun_out/r9_launches.csv | cut -d, -f5,12- | tail -n 16
