#!/bin/bash
# Round-2 GPU pass P (8 GPUs): UJ_fmm (LET) at 5M on 8 / 4 / 2 GPUs after the sparse-leaf refinement, parity vs one GPU at 8.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 240 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 5 --let-timing ) > gpurun_out/p_bench_fmm_5m_8gpu.json 2> gpurun_out/p_bench_fmm_5m_8gpu.err
( time timeout 240 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --uj fmm --particles 5000000 --steps 5 --warmup 5 --no-parity ) > gpurun_out/p_bench_fmm_5m_4gpu.json 2> gpurun_out/p_bench_fmm_5m_4gpu.err
( time timeout 240 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --uj fmm --particles 5000000 --steps 5 --warmup 5 --no-parity ) > gpurun_out/p_bench_fmm_5m_2gpu.json 2> gpurun_out/p_bench_fmm_5m_2gpu.err
( time timeout 240 python bench.py --uj fmm --particles 5000000 --steps 5 --warmup 5 --no-parity ) > gpurun_out/p_bench_fmm_5m_1gpu.json 2> gpurun_out/p_bench_fmm_5m_1gpu.err
