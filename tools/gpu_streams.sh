#!/bin/bash
# sweep of the number of concurrent sub-batch chains of the sampler (cfg2, one box)
mkdir -p gpurun_out
for n in 5 4 6 7 8 10 5; do
  SD_SAMPLER_STREAMS=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/st_$n.log 2>&1
  echo "streams $n: $(tail -n 1 gpurun_out/st_$n.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['e2e']['value'], j['gpu_launches'])")"
done
