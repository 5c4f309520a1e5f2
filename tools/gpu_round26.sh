#!/bin/bash
mkdir -p gpurun_out
SD_TC_PAIR=1 timeout 120 python tools/diag_tc.py > gpurun_out/r26_diag.log 2>&1; echo "rc=$?" >> gpurun_out/r26_diag.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv >> gpurun_out/r26_diag.log 2>&1
SD_TC_PAIR=1 timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "tc" -x > gpurun_out/r26_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r26_tc.log
tail -n 25 gpurun_out/r26_diag.log | cut -c1-250; tail -n 15 gpurun_out/r26_tc.log | cut -c1-250
