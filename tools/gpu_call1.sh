#!/bin/bash
# round-1 re-entry, GPU call 1: parity of the re-written FMM near field + new tables, sweep, ncu of the leaf kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/c1_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/c1_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/c1_tests.log
timeout 120 compute-sanitizer --tool memcheck python tools/prof_fmm.py 20000 > gpurun_out/c1_memcheck.log 2>&1
timeout 200 compute-sanitizer --tool racecheck python tools/prof_fmm.py 20000 > gpurun_out/c1_racecheck.log 2>&1
rm -f gpurun_out/sweep.jsonl
( time timeout 600 python tools/sweep.py --cases rotor,vahana,wing,random --max-n 20000000 ) > gpurun_out/c1_sweep.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmm_leaf -c 2 -f -o gpurun_out/r01b_fmm_leaf python tools/prof_fmm.py 1000000 > gpurun_out/c1_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_fmm_n1m.csv python tools/prof_fmm.py 1000000 > gpurun_out/c1_ncu2.log 2>&1
tail -5 gpurun_out/c1_tests.log; tail -3 gpurun_out/c1_sweep.log
