#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/c5_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/c5_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
tail -4 gpurun_out/c5_tests.log; cut -c1-1500 gpurun_out/c5_bench.json
