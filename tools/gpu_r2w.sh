#!/bin/bash
# Round-2 GPU pass W (2 GPUs): the demand-driven halo LET mode over NCCL — the two let_halo cases of tests/test_gpu_dist.py,
# then the 5M-ring UJ_fmm step in halo mode (with parity against one GPU, memory and halo volume per rank) and, on the same
# box, the all-gather mode.
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q -k let_halo ) > gpurun_out/w_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/w_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 200 $TR --master-port 29511 bench.py --gpus 2 --uj fmm --particles 5000000 --fmm-mode let_halo --let-timing --steps 3 --warmup 3 ) > gpurun_out/w_bench_fmm_5m_halo_2gpu.json 2> gpurun_out/w_bench_fmm_5m_halo_2gpu.err
( time timeout 200 $TR --master-port 29512 bench.py --gpus 2 --uj fmm --particles 5000000 --fmm-mode let --steps 3 --warmup 3 --no-parity ) > gpurun_out/w_bench_fmm_5m_let_2gpu.json 2> gpurun_out/w_bench_fmm_5m_let_2gpu.err
tail -6 gpurun_out/w_tests.log | cut -c1-200; cut -c1-400 gpurun_out/w_bench_fmm_5m_halo_2gpu.json; tail -4 gpurun_out/w_bench_fmm_5m_halo_2gpu.err; cut -c1-400 gpurun_out/w_bench_fmm_5m_let_2gpu.json; tail -4 gpurun_out/w_bench_fmm_5m_let_2gpu.err
