#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/diag_step.py > gpurun_out/c6_diag.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:uj_direct_f64 -c 1 -f -o gpurun_out/r01d_uj_f64_n1m python tools/prof_uj.py 1000000 gaussianerf 1 > gpurun_out/c6_ncu_k1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01d_launches_bench_n200k.csv python bench.py --particles 200000 --steps 2 --warmup 1 --no-cpu --no-e2e --no-fmm > gpurun_out/c6_ncu_bench.log 2>&1
grep -E "engine step|host-buffer" gpurun_out/c6_diag.log | cut -c1-60
tail -3 gpurun_out/c6_ncu_k1.log
