#!/bin/bash
# number of concurrent sub-batch chains of the sampler at shard sizes that fill most of one wave of tile pairs
# (cfg2 shapes, one box, interleaved)
mkdir -p gpurun_out
for rep in 1 2; do for cfg in "224 5" "224 9" "256 5" "256 10" "320 5" "320 10" "384 5" "384 10"; do set -- $cfg
  SD_SAMPLER_STREAMS=$2 timeout 300 python bench.py --workload cfg2 --batch $1 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/st_$1_$2.log 2>&1
  echo "rep $rep b=$1 streams $2: $(tail -n 1 gpurun_out/st_$1_$2.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['e2e']['value'])")"
done; done
