#!/bin/bash
# round-2 artefacts for profiles/: cycle-stamp traces, parity report, full bench lines (ours + reference arm), smoke
mkdir -p gpurun_out
for cfg in "cfg2 52" "cfg3 512" "cfg4 256"; do set -- $cfg; python tools/trace_tc.py $1 $2 3 > gpurun_out/r02_trace_$1.txt 2>&1; done
python tools/trace_tc.py cfg2 52 2 > gpurun_out/r02_trace_cfg2_fp16.txt 2>&1
python tools/parity_report.py > gpurun_out/r02_parity_report.txt 2>&1
python tools/diag_i8.py > gpurun_out/r02_root_flips.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.log 2>&1; echo "ref rc=$?"
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
tail -n 3 gpurun_out/r02_smoke.log; tail -n 1 gpurun_out/r02_bench.log | cut -c1-200
