#!/bin/bash
# Round-2 GPU pass U (1 GPU): final validation — smoke, the whole GPU suite, bench line, random-field FMM baseline.
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/u_smoke.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/u_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/u_tests.log
( time timeout 900 python bench.py ) > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
( time timeout 300 python bench.py --uj fmm --particles 5000000 --field random --steps 3 --warmup 3 --no-parity ) > gpurun_out/u_bench_fmm_5m_random_1gpu.json 2> gpurun_out/u_bench_fmm_5m_random_1gpu.err
grep "smoke" gpurun_out/u_smoke.log; tail -16 gpurun_out/u_tests.log | cut -c1-200; cut -c1-300 gpurun_out/u_bench.json; tail -2 gpurun_out/u_bench.err; cut -c1-330 gpurun_out/u_bench_fmm_5m_random_1gpu.json
