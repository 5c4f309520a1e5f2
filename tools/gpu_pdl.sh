#!/bin/bash
# programmatic dependent launch: large 2-stream shards (cfg2 shapes b=1024, cfg3) with it always on vs off
mkdir -p gpurun_out
for wl in "cfg2 1024" "cfg3 1024" "cfg2 512"; do set -- $wl; for mode in 0 2; do
    SD_PDL=$mode timeout 300 python bench.py --workload $1 --batch $2 --steps 4 --warmup 2 --no-secondary --no-cpu-baseline > gpurun_out/pdl_${mode}_$1_$2.log 2>&1
    echo "SD_PDL=$mode $1 b=$2: $(tail -n 1 gpurun_out/pdl_${mode}_$1_$2.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'])" 2>&1 | tail -n 1)"
done; done
