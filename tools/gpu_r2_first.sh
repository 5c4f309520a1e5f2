#!/bin/bash
# round-2 first call: kind::i8 probe, GPU test suite, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 ./gpurun_out/probe_i8 > gpurun_out/r2a_probe_i8.log 2>&1; echo "probe rc=$?" >> gpurun_out/r2a_probe_i8.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2>&1
cat gpurun_out/r2a_probe_i8.log; tail -n 3 gpurun_out/r2a_tests.log; tail -n 1 gpurun_out/r2a_bench.log | cut -c1-1500
