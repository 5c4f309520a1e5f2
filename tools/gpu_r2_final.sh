#!/bin/bash
# regenerate the round-2 artefacts of profiles/ after a kernel change: traces, parity report, bench lines, launch list,
# full ncu captures of the tcgen05 layers (25 launches of one step) and of conv1, sanitizer
mkdir -p gpurun_out
bash tools/gpu_r2_profiles.sh > gpurun_out/r02_profiles_run.log 2>&1
bash tools/gpu_launch_list.sh p_launches > gpurun_out/r02_launch_run.log 2>&1
bash tools/gpu_ncu_tc25.sh > gpurun_out/r02_ncu25_run.log 2>&1
export SD_SAMPLER_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_real_const -s 10 -c 1 -f -o gpurun_out/p_conv1 python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/p_conv1.log 2>&1
unset SD_SAMPLER_GRAPH
timeout 600 python bench.py --workload ref16 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ref16_b32.log 2>&1
timeout 600 python bench.py --workload ref16 --batch 16 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ref16_b16.log 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/r02_sanitize_run.log 2>&1
tail -n 4 gpurun_out/r02_profiles_run.log | cut -c1-300; tail -n 2 gpurun_out/r02_launch_run.log; grep -E "SUMMARY" gpurun_out/san_*.log
for f in gpurun_out/r02_ref16_b*.log; do tail -n 1 $f | cut -c1-160; done
