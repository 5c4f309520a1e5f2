#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r33_tests.log 2>&1
tail -n 3 gpurun_out/r33_tests.log
run() { # workload batch steps
timeout 600 python bench.py --workload $1 --batch $2 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/r33_$1_b$2.log 2>&1
echo "$1 b=$2: $(tail -n 1 gpurun_out/r33_$1_b$2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['config']['sampler_streams'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['whole_step_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons'])")"
}
run cfg2 256 5; run cfg2 64 5; run cfg2 100 5; run cfg2 128 5; run cfg2 192 5; run cfg2 384 5; run cfg3 1024 3; run cfg4 512 2; run ref16 32 5
SD_SAMPLER_STREAMS=1 timeout 600 python bench.py --workload cfg2 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline | tail -n 1 | cut -c1-140
SD_SAMPLER_STREAMS=3 timeout 600 python bench.py --workload cfg2 --batch 128 --steps 5 --warmup 3 --no-cpu-baseline | tail -n 1 | cut -c1-140
