#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/r3_tests.log 2>&1
timeout 600 python tools/bench_layers.py cfg2 > gpurun_out/r3_layers.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 42 -c 2 -o gpurun_out/r3_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_ncu_full.log 2>&1
for f in gpurun_out/r3_*.log; do echo "=== $f"; tail -n 8 $f | cut -c1-1200; done
