#!/bin/bash
# Round-2 GPU pass D (2 GPUs): multi-GPU handle behind the C ABI, the tests fixed since pass C, LET phase timing at 5M.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_simloop.py tests/test_gpu_let.py -q --durations=8 ) > gpurun_out/d_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/d_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 python bench.py --uj fmm --particles 5000000 --steps 2 --warmup 3 --let-timing --no-parity ) > gpurun_out/d_bench_fmm_5m_1gpu.json 2> gpurun_out/d_bench_fmm_5m_1gpu.err
( time timeout 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --uj fmm --particles 5000000 --steps 2 --warmup 3 --let-timing --no-parity ) > gpurun_out/d_bench_fmm_5m_2gpu.json 2> gpurun_out/d_bench_fmm_5m_2gpu.err
( time timeout 300 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --uj fmm --particles 5000000 --steps 2 --warmup 3 --let-timing --no-parity --fmm-mode replicated ) > gpurun_out/d_bench_fmm_5m_2gpu_repl.json 2> gpurun_out/d_bench_fmm_5m_2gpu_repl.err
tail -12 gpurun_out/d_tests.log; for f in d_bench_fmm_5m_1gpu d_bench_fmm_5m_2gpu d_bench_fmm_5m_2gpu_repl; do echo "== $f"; cut -c1-1800 gpurun_out/$f.json; tail -3 gpurun_out/$f.err | cut -c1-300; done
