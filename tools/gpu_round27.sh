#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py cfg2 > gpurun_out/r27_layers_cfg2.log 2>&1
timeout 300 python tools/bench_layers.py cfg3 > gpurun_out/r27_layers_cfg3.log 2>&1
SD_TC_PAIR=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r27_bench_pair.log 2>&1
SD_TC_PAIR=1 SD_SAMPLER_STREAMS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r27_bench_pair_s1.log 2>&1
SD_TC_PAIR=1 timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r27_cfg3_pair.log 2>&1
for f in gpurun_out/r27_layers*.log; do echo "=== $f"; cat $f | cut -c1-600; done
for f in gpurun_out/r27_bench*.log gpurun_out/r27_cfg3*.log; do echo "=== $f"; tail -n 1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['roofline']['kernel'], d['roofline']['achieved'], {k:v['ms'] for k,v in d['roofline']['layers'].items()}, d['roofline']['whole_step_tflops'])"; done
