#!/bin/bash
# the as-shipped configuration (T=16, 16 / 32 images): parity tests, throughput old vs new library, launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_sampling.py tests/test_gpu_sample_parity.py -m gpu -x -q > gpurun_out/r16_tests.log 2>&1; tail -n 2 gpurun_out/r16_tests.log
for lib in old new; do for b in 16 32; do
  if [ $lib = old ]; then export SD_B200_LIB=build/libsd_b200_old.so; else unset SD_B200_LIB; fi
  timeout 300 python bench.py --workload ref16 --batch $b --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r16_${lib}_b$b.log 2>&1
  echo "$lib b=$b: $(tail -n 1 gpurun_out/r16_${lib}_b$b.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['gpu_launches'])")"
done; done
unset SD_B200_LIB
bash tools/gpu_ab.sh | cut -c1-90
export SD_SAMPLER_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 24 --csv --log-file gpurun_out/r16_launches.csv python bench.py --workload ref16 --batch 16 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r16_list.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r16_launches.csv')) if len(r)>8]
h=rows[0]; ki,vi,ui,gi=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit'),h.index('Grid Size')
for r in rows[1:14]:
    print(r[ki].split('(')[0][-45:], r[gi], r[vi], r[ui])
PY
