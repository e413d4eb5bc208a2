#!/bin/bash
# Round-2 GPU pass C (2 GPUs): NCCL tests of the sharded field (direct + LET + replicated), the tests that failed in pass B,
# a short sharded direct bench with its parity key, and UJ_fmm at 5M on 1 and 2 GPUs (LET).
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
( time timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_let.py tests/test_gpu_simloop.py "tests/test_gpu_scale.py::test_tile_skip_fires_on_wake_fields" -q --durations=8 ) > gpurun_out/c_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/c_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --particles 400000 ) > gpurun_out/c_bench_direct_2gpu.json 2> gpurun_out/c_bench_direct_2gpu.err
( time timeout 300 python bench.py --uj fmm --particles 5000000 --steps 3 --warmup 3 ) > gpurun_out/c_bench_fmm_5m_1gpu.json 2> gpurun_out/c_bench_fmm_5m_1gpu.err
( time timeout 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --uj fmm --particles 5000000 --steps 3 --warmup 3 ) > gpurun_out/c_bench_fmm_5m_2gpu.json 2> gpurun_out/c_bench_fmm_5m_2gpu.err
( time timeout 300 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 ) > gpurun_out/c_bench_ref_2gpu.json 2> gpurun_out/c_bench_ref_2gpu.err
tail -15 gpurun_out/c_tests.log; for f in c_bench_direct_2gpu c_bench_fmm_5m_1gpu c_bench_fmm_5m_2gpu c_bench_ref_2gpu; do echo "== $f"; cut -c1-900 gpurun_out/$f.json; tail -4 gpurun_out/$f.err | cut -c1-300; done
