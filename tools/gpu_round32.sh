#!/bin/bash
mkdir -p gpurun_out
run() { # workload batch streams steps
SD_SAMPLER_STREAMS=$3 timeout 600 python bench.py --workload $1 --batch $2 --steps $4 --warmup 3 --no-cpu-baseline > gpurun_out/r32_$1_b$2_s$3.log 2>&1
echo "$1 b=$2 streams=$3: $(tail -n 1 gpurun_out/r32_$1_b$2_s$3.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['roofline']['whole_step_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons'])")"
}
for s in 5 6 7 8 10; do run cfg2 256 $s 5; done
for s in 1 2 3 4; do run cfg2 128 $s 5; done
for s in 2 3 5 8 10; do run cfg2 512 $s 4; done
for s in 2 4 8; do run cfg4 512 $s 2; done
