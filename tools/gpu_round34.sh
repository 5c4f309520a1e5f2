#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r34_tests.log 2>&1
tail -n 3 gpurun_out/r34_tests.log
for tc in 0 1; do
echo "## SD_DECODER_TC=$tc"
SD_DECODER_TC=$tc timeout 300 python tools/bench_vqvae.py 64 4 2>&1 | tail -1
SD_DECODER_TC=$tc timeout 300 python tools/bench_vqvae.py 1024 4 2>&1 | tail -1
SD_DECODER_TC=$tc timeout 300 python tools/bench_vqvae.py 64 16 2>&1 | tail -1
SD_DECODER_TC=$tc timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
