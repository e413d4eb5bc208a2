#!/bin/bash
# Round-2 GPU pass T (1 GPU): multi-handle FMM, FMM error tests at BASELINE sizes, error table with the box-gap clearance.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_let.py "tests/test_gpu_scale.py::test_fmm_error_vs_direct_at_baseline_sizes" tests/test_gpu_fmm.py -q --durations=6 ) > gpurun_out/t_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/t_tests.log
tail -25 gpurun_out/t_tests.log | cut -c1-220
( time timeout 500 python tools/fmm_error_table.py ) > gpurun_out/t_fmm_error_table.md 2> gpurun_out/t_fmm_error_table.err
cat gpurun_out/t_fmm_error_table.md; tail -2 gpurun_out/t_fmm_error_table.err
