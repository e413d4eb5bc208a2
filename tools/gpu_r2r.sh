#!/bin/bash
# Round-2 GPU pass R (1 GPU): CUDA FMM vs the FMM oracle, the FMM error table, ncu --set full of the FMM kernels at N = 1M.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fmm.py -q --durations=5 ) > gpurun_out/r_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/r_tests.log
tail -25 gpurun_out/r_tests.log
( time timeout 600 python tools/fmm_error_table.py ) > gpurun_out/r_fmm_error_table.md 2> gpurun_out/r_fmm_error_table.err
cat gpurun_out/r_fmm_error_table.md; tail -3 gpurun_out/r_fmm_error_table.err
for k in fmm_m2l_kernel fmm_traverse_kernel fmm_leaf_estr_kernel fmm_leaf_uj_kernel fmm_p2m_kernel; do
  skip=1; [ $k = fmm_traverse_kernel ] && skip=20
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02r_$k python tools/prof_fmm.py 1000000 > gpurun_out/r_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02r_launches_fmm_n1m.csv python tools/prof_fmm.py 1000000 > gpurun_out/r_ncu_launches.log 2>&1
ls -la gpurun_out/r02r_* | cut -c30-
