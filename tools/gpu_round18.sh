#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q > gpurun_out/r18_train.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_training.py > gpurun_out/r18_tests.log 2>&1
tail -n 40 gpurun_out/r18_train.log | cut -c1-300; tail -n 5 gpurun_out/r18_tests.log
