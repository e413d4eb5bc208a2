#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python tools/let_balance_probe.py 5000000 8 rings 11 ) > gpurun_out/n_probe_rings_evolved.log 2>&1
grep -v "^rank" gpurun_out/n_probe_rings_evolved.log | tail -12; grep "^rank" gpurun_out/n_probe_rings_evolved.log | tail -8 | cut -c1-260
( time timeout 600 python -m pytest tests/test_gpu_fmm.py tests/test_gpu_let.py tests/test_gpu_viscous.py -q --durations=5 ) > gpurun_out/n_tests.log 2>&1
tail -15 gpurun_out/n_tests.log
( time timeout 300 python bench.py --uj fmm --particles 5000000 --steps 5 --warmup 11 --no-parity ) > gpurun_out/n_bench_fmm_5m_1gpu.json 2> gpurun_out/n_bench_fmm_5m_1gpu.err
cut -c1-420 gpurun_out/n_bench_fmm_5m_1gpu.json
