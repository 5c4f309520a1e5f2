#!/bin/bash
# epilogue change: conv + model parity tests, cycle-stamp trace, same-box A/B against build/libsd_b200_old.so
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/epi_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/epi_tests.log
timeout 300 python tools/trace_tc.py cfg2 256 3 > gpurun_out/epi_trace.log 2>&1
bash tools/gpu_ab.sh > gpurun_out/epi_ab.log 2>&1
tail -n 3 gpurun_out/epi_tests.log
grep -E "^== conv[2345]|pass 1" gpurun_out/epi_trace.log | sed -E 's/.*(epilogue [0-9]+.*)/\1/' | cut -c1-330
cat gpurun_out/epi_ab.log
