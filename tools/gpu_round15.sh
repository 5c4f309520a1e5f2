#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1
tail -n 80 gpurun_out/parity_report.txt | cut -c1-260
