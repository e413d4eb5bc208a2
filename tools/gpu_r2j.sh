#!/bin/bash
# Round-2 GPU pass J (8 GPUs): LET step at 5M with the work-weighted cut: per-rank table, parity vs one GPU.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 240 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 5 --let-timing ) > gpurun_out/j_bench_fmm_5m_8gpu.json 2> gpurun_out/j_bench_fmm_5m_8gpu.err
( time timeout 240 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --uj fmm --particles 5000000 --steps 5 --warmup 5 --let-timing --no-parity ) > gpurun_out/j_bench_fmm_5m_4gpu.json 2> gpurun_out/j_bench_fmm_5m_4gpu.err
