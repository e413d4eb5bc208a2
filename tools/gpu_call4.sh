#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fmm.py tests/test_gpu_step.py -x -q ) > gpurun_out/c4_tests.log 2>&1
timeout 120 compute-sanitizer --tool memcheck python tools/prof_fmm.py 20000 0 8 > gpurun_out/c4_memcheck.log 2>&1
timeout 200 compute-sanitizer --tool racecheck python tools/prof_fmm.py 20000 0 8 > gpurun_out/c4_racecheck.log 2>&1
for c in 8 1; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_fmm_n1m_copies$c.csv python tools/prof_fmm.py 1000000 0 $c > gpurun_out/c4_ncu_$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmm_leaf_uj -c 1 -f -o gpurun_out/r01d_fmm_leaf_copies8 python tools/prof_fmm.py 1000000 0 8 > gpurun_out/c4_ncu_full.log 2>&1
tail -3 gpurun_out/c4_tests.log; tail -2 gpurun_out/c4_memcheck.log gpurun_out/c4_racecheck.log
grep -h "fmm_leaf" gpurun_out/r01d_launches_fmm_n1m_copies8.csv | head -4
grep -h "fmm_leaf" gpurun_out/r01d_launches_fmm_n1m_copies1.csv | head -4
