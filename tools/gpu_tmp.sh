#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_sampling.py -m gpu -x -q > gpurun_out/bp_tests.log 2>&1; tail -n 2 gpurun_out/bp_tests.log
for lib in old new; do for b in 16 32; do
  if [ $lib = old ]; then export SD_B200_LIB=build/libsd_b200_old.so; else unset SD_B200_LIB; fi
  timeout 300 python bench.py --workload ref16 --batch $b --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r16_${lib}_b$b.log 2>&1
  echo "$lib b=$b: $(tail -n 1 gpurun_out/r16_${lib}_b$b.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'])")"
done; done
unset SD_B200_LIB
bash tools/gpu_ab.sh | cut -c1-200
python tools/trace_tc.py cfg2 52 3 2>&1 | grep -E "^== |pass 1:" | cut -c1-150
