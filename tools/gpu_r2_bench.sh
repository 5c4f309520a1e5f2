#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sampling.py -x -q > gpurun_out/r2b_tests.log 2>&1; tail -n 3 gpurun_out/r2b_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/r2b_bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_ref.log 2>&1; echo "ref rc=$?"
tail -n 1 gpurun_out/r2b_ref.log | cut -c1-300
