#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r17_tests.log 2>&1
timeout 900 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1
tail -n 5 gpurun_out/r17_tests.log; grep -E "^##|^den|^enc|^gen|^dec|logits|VQ|recon" gpurun_out/parity_report.txt | cut -c1-110
