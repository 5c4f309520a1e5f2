#!/bin/bash
mkdir -p gpurun_out
export SD_SAMPLER_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lif_from_currents -s 20 -c 2 -f -o gpurun_out/p_lifcur python bench.py --workload ref16 --batch 16 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/p_lifcur.log 2>&1
echo rc=$?
