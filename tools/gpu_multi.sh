#!/bin/bash
# N-GPU bench line exactly as the driver launches it (one rank per GPU, torchrun, NCCL); usage: gpu_multi.sh N [steps]
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt 2>&1
N=${1:-2}
S=${2:-5}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $S --warmup 3 > gpurun_out/m_bench_$N.log 2>&1
echo "rc=$?" >> gpurun_out/m_bench_$N.log
tail -n 3 gpurun_out/m_bench_$N.log | cut -c1-400
