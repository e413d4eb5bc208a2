#!/bin/bash
# Round-2 GPU pass B (1 GPU): the local-essential-tree tests (loopback collectives), the at-scale parity tests, then the whole suite.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_let.py -q --durations=8 -x ) > gpurun_out/b_let.log 2>&1
echo "let exit: $?" >> gpurun_out/b_let.log
( time timeout 1200 python -m pytest tests -m gpu -q --durations=12 --deselect tests/test_gpu_let.py ) > gpurun_out/b_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/b_tests.log
tail -40 gpurun_out/b_let.log; tail -25 gpurun_out/b_tests.log
