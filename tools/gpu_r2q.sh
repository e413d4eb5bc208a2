#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 240 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --uj fmm --particles 5000000 --steps 5 --warmup 5 --no-parity --let-timing ) > gpurun_out/q_bench_fmm_5m_2gpu.json 2> gpurun_out/q_bench_fmm_5m_2gpu.err
( time timeout 240 $TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --uj fmm --particles 5000000 --steps 5 --warmup 5 --no-parity --let-timing --no-balance ) > gpurun_out/q_bench_fmm_5m_2gpu_nobal.json 2> gpurun_out/q_bench_fmm_5m_2gpu_nobal.err
