#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt 2>&1
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/m_bench_$N.log 2>&1
echo "rc=$?" >> gpurun_out/m_bench_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/m_ref_$N.log 2>&1
echo "rc=$?" >> gpurun_out/m_ref_$N.log
tail -n 5 gpurun_out/m_bench_$N.log | cut -c1-1200; tail -n 3 gpurun_out/m_ref_$N.log | cut -c1-600
