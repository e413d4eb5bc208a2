#!/bin/bash
# Round-2 GPU pass G (8 GPUs): UJ_fmm (LET) at 5M on 8 GPUs with the work-weighted cut on / off, parity vs one GPU; LET tests.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
( time timeout 300 python -m pytest tests/test_gpu_let.py -q -x ) > gpurun_out/g_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/g_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 240 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 5 --let-timing ) > gpurun_out/g_bench_fmm_5m_8gpu.json 2> gpurun_out/g_bench_fmm_5m_8gpu.err
( time timeout 240 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 5 --let-timing --no-balance --no-parity ) > gpurun_out/g_bench_fmm_5m_8gpu_nobal.json 2> gpurun_out/g_bench_fmm_5m_8gpu_nobal.err
tail -4 gpurun_out/g_tests.log; for f in g_bench_fmm_5m_8gpu g_bench_fmm_5m_8gpu_nobal; do echo "== $f"; cut -c1-300 gpurun_out/$f.json; tail -3 gpurun_out/$f.err | cut -c1-300; done
