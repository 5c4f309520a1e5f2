#!/bin/bash
# launch list of one bench step (un-graphed sampler; ncu serialises launches: compare SHARES, not absolutes)
mkdir -p gpurun_out
export SD_SAMPLER_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1990 -c 2100 --csv --log-file gpurun_out/${1:-r2_launches}.csv python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/${1:-r2_launches}.log 2>&1
echo "rc=$?"; wc -l gpurun_out/${1:-r2_launches}.csv
