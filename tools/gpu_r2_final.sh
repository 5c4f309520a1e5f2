#!/bin/bash
# regenerate the round-2 artefacts of profiles/ after a kernel change: traces, parity report, bench lines, launch list, sanitizer
mkdir -p gpurun_out
bash tools/gpu_r2_profiles.sh > gpurun_out/r02_profiles_run.log 2>&1
bash tools/gpu_launch_list.sh r2h_launches > gpurun_out/r02_launch_run.log 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/r02_sanitize_run.log 2>&1
tail -n 4 gpurun_out/r02_profiles_run.log | cut -c1-300; tail -n 2 gpurun_out/r02_launch_run.log; grep -E "SUMMARY" gpurun_out/san_*.log
