#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_train.py > gpurun_out/r20.log 2>&1
cat gpurun_out/r20.log | cut -c1-200
