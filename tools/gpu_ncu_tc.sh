#!/bin/bash
# one full ncu capture of the five tcgen05 layers of one sub-batch (launch order conv2..conv6), un-graphed sampler;
# dense warp-state sampling (--warp-sampling-interval 0) so that the epilogue's ~7 k cycles per pass get enough samples
mkdir -p gpurun_out
export SD_SAMPLER_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:conv3x3_tc -s 50 -c 5 -f -o gpurun_out/${1:-r2_conv_tc} python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/${1:-r2_conv_tc}.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${1:-r2_conv_tc}.ncu-rep
