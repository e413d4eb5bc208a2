#!/bin/bash
# Round-2 GPU pass S (1 GPU): the whole validation — smoke, the full GPU test suite, both bench arms, ncu summaries of the FMM
# kernels at N = 1M (reports are summarised on the box and deleted: gpurun_out/ must stay under 64 MiB).
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/s_smoke.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/s_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/s_tests.log
( time timeout 900 python bench.py ) > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/s_bench_reference.json 2> gpurun_out/s_bench_reference.err
for k in fmm_m2l_kernel fmm_traverse_kernel fmm_leaf_estr_kernel fmm_leaf_uj_kernel; do
  skip=1; [ $k = fmm_traverse_kernel ] && skip=20
  timeout 300 ncu --set full --clock-control none -k regex:$k -s $skip -c 1 -f -o /tmp/r02s_$k python tools/prof_fmm.py 1000000 > gpurun_out/s_ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/r02s_$k.ncu-rep gpurun_out/r02s_$k.txt "$k, N = 1,000,000 vortex rings, UJ_fmm p=4 ncrit=50 theta=0.4 (round 2)" > /dev/null 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s_launches_fmm_n1m.csv python tools/prof_fmm.py 1000000 > gpurun_out/s_ncu_launches.log 2>&1
tail -2 gpurun_out/s_smoke.log; tail -22 gpurun_out/s_tests.log | cut -c1-200; cut -c1-600 gpurun_out/s_bench.json; tail -2 gpurun_out/s_bench.err; cut -c1-300 gpurun_out/s_bench_reference.json; du -sh gpurun_out
