#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r31_tests.log 2>&1
tail -n 3 gpurun_out/r31_tests.log
run() { # workload streams steps
SD_SAMPLER_STREAMS=$2 timeout 600 python bench.py --workload $1 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/r31_$1_s$2.log 2>&1
echo "$1 streams=$2: $(tail -n 1 gpurun_out/r31_$1_s$2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['whole_step_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons'])")"
}
for s in 2 3 4 5; do run cfg2 $s 5; done
for s in 2 3 4 6 8; do run cfg3 $s 3; done
for s in 2 3 4; do run cfg4 $s 2; done
run ref16 1 5
