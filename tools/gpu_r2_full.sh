#!/bin/bash
# full GPU check after a kernel change: whole GPU test suite, tcgen05 fuzz, default bench (with the secondary records)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/full_tests.log
timeout 600 python tools/fuzz_tc.py 60 3 > gpurun_out/full_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/full_fuzz.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/full_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/full_bench.log
tail -n 3 gpurun_out/full_tests.log; tail -n 4 gpurun_out/full_fuzz.log; tail -n 2 gpurun_out/full_bench.log | cut -c1-600
