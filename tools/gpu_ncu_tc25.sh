#!/bin/bash
# full ncu capture of one diffusion step's tcgen05 launches: 5 layers x 5 sub-batches, in launch order (sub-batch major)
mkdir -p gpurun_out
export SD_SAMPLER_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 50 -c 25 -f -o gpurun_out/p_conv_tc python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/p_tc.log 2>&1
echo "rc=$?"; ls -la gpurun_out/p_conv_tc.ncu-rep
