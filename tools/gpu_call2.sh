#!/bin/bash
# round-1 re-entry, GPU call 2: W=0.5/deg-9 G table + persistent leaf warps; full tests, sweep (all configs), ncu, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/c2_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/c2_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmm_leaf -c 2 -f -o gpurun_out/r01c_fmm_leaf python tools/prof_fmm.py 1000000 > gpurun_out/c2_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches_fmm_n1m.csv python tools/prof_fmm.py 1000000 > gpurun_out/c2_ncu2.log 2>&1
rm -f gpurun_out/sweep.jsonl
( time timeout 900 python tools/sweep.py --cases rotor,vahana,wing,random --max-n 50000000 ) > gpurun_out/c2_sweep.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
tail -4 gpurun_out/c2_tests.log; tail -2 gpurun_out/c2_sweep.log; tail -c 600 gpurun_out/c2_bench.json
