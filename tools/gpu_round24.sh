#!/bin/bash
mkdir -p gpurun_out
SD_TC_WIDE256=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r24_wide.log 2>&1
SD_TC_WIDE256=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r24_base.log 2>&1
SD_TC_WIDE256=1 SD_SAMPLER_STREAMS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r24_wide_s1.log 2>&1
SD_TC_WIDE256=1 SD_SAMPLER_STREAMS=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r24_wide_s3.log 2>&1
for f in gpurun_out/r24_*.log; do echo "=== $f"; tail -n 1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['roofline']['kernel'], d['roofline']['achieved'], {k:v['ms'] for k,v in d['roofline']['layers'].items()}, d['roofline']['whole_step_tflops'])"; done
