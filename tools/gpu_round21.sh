#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r21_tests.log 2>&1
tail -n 6 gpurun_out/r21_tests.log | cut -c1-300
