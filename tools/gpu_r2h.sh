#!/bin/bash
# Round-2 GPU pass H (8 GPUs): per-rank phase table of the LET step at 5M (count-based cut).
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 240 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 3 --warmup 3 --let-timing --no-balance --no-parity ) > gpurun_out/h_bench_fmm_5m_8gpu_nobal.json 2> gpurun_out/h_bench_fmm_5m_8gpu_nobal.err
( time timeout 240 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 3 --warmup 3 --let-timing --no-balance --no-parity --field random ) > gpurun_out/h_bench_fmm_5m_8gpu_random.json 2> gpurun_out/h_bench_fmm_5m_8gpu_random.err
nproc > gpurun_out/h_nproc.txt; nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader >> gpurun_out/h_nproc.txt
