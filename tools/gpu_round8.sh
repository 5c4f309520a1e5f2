#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_lif.py > gpurun_out/r8_lif.log 2>&1
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r8_cfg3.log 2>&1
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r8_cfg4.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r8_ref.log 2>&1
for f in gpurun_out/r8_*.log; do echo "=== $f"; tail -n 30 $f | cut -c1-1500; done
