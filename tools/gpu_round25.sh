#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_train.py > gpurun_out/r25_train.log 2>&1
tail -n 25 gpurun_out/r25_train.log
