#!/bin/bash
# full ncu captures of the tcgen05 layers of one diffusion step at cfg3 (b=1024, 8x8 grid) and cfg4 (T=8, K=512, b=512):
# 5 layers x 2 sub-batch chains each, un-graphed sampler
mkdir -p gpurun_out
export SD_SAMPLER_GRAPH=0
timeout 900 ncu --set full --clock-control none -k regex:conv3x3_tc -s 20 -c 10 -f -o gpurun_out/p_conv_tc_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/p_tc_cfg3.log 2>&1; echo "cfg3 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:conv3x3_tc -s 20 -c 10 -f -o gpurun_out/p_conv_tc_cfg4 python bench.py --workload cfg4 --batch 512 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/p_tc_cfg4.log 2>&1; echo "cfg4 rc=$?"
ls -la gpurun_out/p_conv_tc_cfg*.ncu-rep
