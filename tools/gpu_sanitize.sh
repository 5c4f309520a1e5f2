#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/san_memcheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize.py > gpurun_out/san_racecheck.log 2>&1
timeout 1200 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize.py > gpurun_out/san_synccheck.log 2>&1
for f in gpurun_out/san_*.log; do echo "=== $f"; tail -n 12 $f | cut -c1-400; done
