#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r7_tests.log 2>&1
for s in 1 2 4 8; do
SD_SAMPLER_STREAMS=$s timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_s$s.log 2>&1
done
SD_SAMPLER_GRAPH=0 SD_SAMPLER_STREAMS=4 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_s4_nograph.log 2>&1
for f in gpurun_out/r7_*.log; do echo "=== $f"; tail -n 5 $f | cut -c1-420; done
