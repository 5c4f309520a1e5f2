#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r28_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r28_smoke.log 2>&1
for s in 2 3 4; do SD_SAMPLER_STREAMS=$s timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r28_bench_s$s.log 2>&1; done
timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r28_cfg4.log 2>&1
timeout 600 python bench.py --workload ref16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r28_ref16.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize.py > gpurun_out/r28_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize.py > gpurun_out/r28_racecheck.log 2>&1
tail -n 4 gpurun_out/r28_tests.log; tail -n 2 gpurun_out/r28_smoke.log; tail -n 2 gpurun_out/r28_memcheck.log gpurun_out/r28_racecheck.log
for f in gpurun_out/r28_bench*.log gpurun_out/r28_cfg4.log gpurun_out/r28_ref16.log; do echo "=== $f"; tail -n 1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['achieved'], {k:v['ms'] for k,v in d['roofline']['layers'].items()}, d['roofline']['whole_step_tflops'], d['clocks'])"; done
