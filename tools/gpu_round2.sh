#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_sampling.py tests/test_gpu_models.py -m gpu -q > gpurun_out/r2_tests.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n64.log 2>&1
SD_TC_NTILE=128 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n128.log 2>&1
SD_TC_KBLK=64 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_k64.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --nsplit 1 > gpurun_out/r2_bench_ns1.log 2>&1
# launch list of one bench run + full profile of the conv4-shaped launch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 4 -o gpurun_out/r2_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
for f in gpurun_out/r2_*.log; do echo "=== $f"; tail -n 6 $f | cut -c1-1500; done
