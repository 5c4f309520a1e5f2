#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r29_tests.log 2>&1
tail -n 3 gpurun_out/r29_tests.log
timeout 300 python tools/trace_tc.py cfg2 256 2>&1 | tail -30
for s in 2; do SD_SAMPLER_STREAMS=$s timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_s$s.log 2>&1; done
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r29_cfg3.log 2>&1
for f in gpurun_out/r29_bench*.log gpurun_out/r29_cfg3.log; do echo "=== $f"; tail -n 1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['achieved'], {k:v['ms'] for k,v in d['roofline']['layers'].items()}, d['roofline']['whole_step_tflops'], d['clocks'])"; done
