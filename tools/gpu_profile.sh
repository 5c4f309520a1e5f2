#!/bin/bash
# Profile artefacts for profiles/: launch list of one bench command + full captures of the top kernels.
# The sampler runs un-graphed here (ncu serialises kernels anyway); sub-batch plan and kernels are the bench defaults.
mkdir -p gpurun_out
export SD_SAMPLER_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1990 -c 2100 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_list.log 2>&1
# one diffusion step of every sub-batch: 5 tcgen05 layers x 5 sub-batches, in launch order (sub-batch major)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 50 -c 25 -o gpurun_out/p_conv_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_step -s 10 -c 1 -o gpurun_out/p_sample python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_sample.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_real_const -s 10 -c 1 -o gpurun_out/p_conv1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_conv1.log 2>&1
unset SD_SAMPLER_GRAPH
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/p_bench.log 2>&1
tail -n 2 gpurun_out/p_bench.log | cut -c1-3000
ls -la gpurun_out/p_*
