#!/bin/bash
mkdir -p gpurun_out
SD_TRAIN_TC=0 timeout 600 python tools/bench_train.py > gpurun_out/train_tc0.log 2>&1
SD_TRAIN_TC=1 timeout 600 python tools/bench_train.py > gpurun_out/train_tc1.log 2>&1
echo "== SD_TRAIN_TC=0"; grep -E "denoiser|vqvae" gpurun_out/train_tc0.log -A1 | grep -v "^--" | cut -c1-200
echo "== SD_TRAIN_TC=1"; grep -E "denoiser|vqvae" gpurun_out/train_tc1.log -A1 | grep -v "^--" | cut -c1-200
