#!/bin/bash
mkdir -p gpurun_out
export SD_TC_SMALL_BATCH_SPLIT=0
for pz in 1 0; do for s in 2 3 4 6; do
SD_TC_PERSIST=$pz SD_SAMPLER_STREAMS=$s timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r30_p${pz}_s$s.log 2>&1
echo "persist=$pz streams=$s: $(tail -n 1 gpurun_out/r30_p${pz}_s$s.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['roofline']['whole_step_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons'])")"
done; done
