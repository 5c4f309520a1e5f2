#!/bin/bash
# every random-configuration cross-check of tools/ on the code as shipped (outputs go to profiles/r02_fuzz_*.txt)
mkdir -p gpurun_out
for f in fused_nonsquare sample tc train_conv train_lif_bn vq vqvae; do
  timeout 600 python tools/fuzz_$f.py > gpurun_out/fz_$f.log 2>&1; echo "fuzz_$f rc=$?"; tail -n 2 gpurun_out/fz_$f.log | cut -c1-200
done
