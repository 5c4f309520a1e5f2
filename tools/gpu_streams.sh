#!/bin/bash
# number of concurrent sub-batch chains of the sampler at intermediate shard sizes (cfg2 shapes, one box, interleaved)
mkdir -p gpurun_out
for rep in 1 2; do for cfg in "64 2" "64 3" "96 3" "96 4" "160 5" "160 7" "192 5" "192 8" "128 5" "128 6"; do set -- $cfg
  SD_SAMPLER_STREAMS=$2 timeout 300 python bench.py --workload cfg2 --batch $1 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/st_$1_$2.log 2>&1
  echo "rep $rep b=$1 streams $2: $(tail -n 1 gpurun_out/st_$1_$2.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['e2e']['value'])")"
done; done
