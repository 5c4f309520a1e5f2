#!/bin/bash
# same-box A/B of two builds of the library: build/libsd_b200_old.so (previous commit) against the in-tree one
mkdir -p gpurun_out
for rep in 1 2; do
  SD_B200_LIB=build/libsd_b200_old.so timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/ab_old_$rep.log 2>&1
  timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/ab_new_$rep.log 2>&1
done
for f in gpurun_out/ab_*.log; do echo "$f $(tail -n 1 $f | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip()); r=j['roofline']
print(j['value'], j['e2e']['value'], r['frac'], r['whole_step']['frac'], {k:v['ms'] for k,v in r['layers'].items()})")"; done
