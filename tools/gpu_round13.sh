#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r13_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r13_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r13_bench.log 2>&1
for f in gpurun_out/r13_*.log; do echo "=== $f"; tail -n 8 $f | cut -c1-2600; done
