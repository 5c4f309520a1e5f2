#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_dec1.py > gpurun_out/r16_a.log 2>&1
SD_SIMT_GENERIC=1 timeout 600 python tools/diag_dec1.py > gpurun_out/r16_b.log 2>&1
cat gpurun_out/r16_a.log gpurun_out/r16_b.log | cut -c1-400
