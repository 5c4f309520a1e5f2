#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q > gpurun_out/r19_train.log 2>&1
tail -n 30 gpurun_out/r19_train.log | cut -c1-300
