#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_let.py tests/test_gpu_fmm.py -q -x ) > gpurun_out/o_tests.log 2>&1
tail -4 gpurun_out/o_tests.log
