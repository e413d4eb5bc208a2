#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fmm.py tests/test_gpu_step.py tests/test_gpu_simloop.py -x -q ) > gpurun_out/c3_tests.log 2>&1
timeout 200 python tools/prof_step.py 1000000 rings 3 > gpurun_out/c3_step_rings.log 2>&1
timeout 200 python tools/prof_step.py 1000000 rotor 3 > gpurun_out/c3_step_rotor.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01c_launches_step_rings.csv python tools/prof_step.py 1000000 rings 1 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_tests.log; cat gpurun_out/c3_step_rings.log gpurun_out/c3_step_rotor.log
