#!/bin/bash
# Round-2 GPU pass F (8 GPUs): NCCL + multi-handle tests on real devices, UJ_fmm (LET) at 5M on 4 and 8 GPUs with phase timing
# and one-GPU parity, the replicated scheme for comparison.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/f_gpus.txt
( time timeout 400 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -q --durations=5 ) > gpurun_out/f_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/f_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 240 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 3 --let-timing ) > gpurun_out/f_bench_fmm_5m_8gpu.json 2> gpurun_out/f_bench_fmm_5m_8gpu.err
( time timeout 240 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --uj fmm --particles 5000000 --steps 5 --warmup 3 --let-timing --no-parity ) > gpurun_out/f_bench_fmm_5m_4gpu.json 2> gpurun_out/f_bench_fmm_5m_4gpu.err
( time timeout 240 python bench.py --uj fmm --particles 5000000 --steps 5 --warmup 3 --no-parity ) > gpurun_out/f_bench_fmm_5m_1gpu.json 2> gpurun_out/f_bench_fmm_5m_1gpu.err
( time timeout 240 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 3 --no-parity --fmm-mode replicated ) > gpurun_out/f_bench_fmm_5m_8gpu_repl.json 2> gpurun_out/f_bench_fmm_5m_8gpu_repl.err
tail -8 gpurun_out/f_tests.log; for f in f_bench_fmm_5m_8gpu f_bench_fmm_5m_4gpu f_bench_fmm_5m_1gpu f_bench_fmm_5m_8gpu_repl; do echo "== $f"; cut -c1-400 gpurun_out/$f.json; tail -3 gpurun_out/$f.err | cut -c1-300; done
