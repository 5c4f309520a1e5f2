#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_layers.py cfg2 > gpurun_out/r11_layers_cfg2.log 2>&1
timeout 600 python tools/bench_layers.py cfg3 > gpurun_out/r11_layers_cfg3.log 2>&1
for f in gpurun_out/r11_*.log; do echo "=== $f"; tail -n 12 $f | cut -c1-900; done
