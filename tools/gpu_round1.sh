#!/bin/bash
# First GPU pass: diagnostics + tests in independent chunks (a hang in one chunk must not hide the others).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 300 python tools/diag_tc.py > gpurun_out/diag_tc.log 2>&1; echo "diag rc=$?" >> gpurun_out/diag_tc.log
timeout 600 python -m pytest tests/test_gpu_lif.py tests/test_gpu_vq.py -m gpu -q > gpurun_out/t1_lif_vq.log 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "simt or unfused" > gpurun_out/t2_simt.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sampling.py -m gpu -q -k "uniform or categorical" > gpurun_out/t4_philox.log 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "tc" > gpurun_out/t3_tc.log 2>&1
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q > gpurun_out/t5_models.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sampling.py -m gpu -q -k "not uniform and not categorical" > gpurun_out/t6_sample.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
for f in gpurun_out/*.log; do echo "=== $f"; tail -n 12 $f; done
