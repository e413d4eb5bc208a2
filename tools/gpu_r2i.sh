#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_let.py tests/test_gpu_multi.py -q -x ) > gpurun_out/i_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/i_tests.log
tail -5 gpurun_out/i_tests.log
