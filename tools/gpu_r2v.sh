#!/bin/bash
# Round-2 GPU pass V (1 GPU): the demand-driven halo LET mode (loopback ranks) + every FMM-touching test after the FmmHalo
# kernel plumbing, and the 5M one-GPU FMM line as the no-regression check of the near-field / M2L kernels.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_let.py tests/test_gpu_fmm.py tests/test_gpu_multi.py tests/test_gpu_simloop.py tests/test_gpu_dist.py -m gpu -q --durations=6 ) > gpurun_out/v_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/v_tests.log
( time timeout 200 python bench.py --uj fmm --particles 5000000 --field random --steps 3 --warmup 3 --no-parity ) > gpurun_out/v_bench_fmm_5m_random_1gpu.json 2> gpurun_out/v_bench_fmm_5m_random_1gpu.err
tail -30 gpurun_out/v_tests.log | cut -c1-220; cut -c1-330 gpurun_out/v_bench_fmm_5m_random_1gpu.json; tail -3 gpurun_out/v_bench_fmm_5m_random_1gpu.err
