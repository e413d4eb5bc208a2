#!/bin/bash
# Round-2 GPU pass A (1 GPU): tests, bench (both arms), launch list of a bench step, ncu --set full of K1 / K2 at N = 1M and
# of the DFMA microbenchmark.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/a_smoke.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/a_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/a_tests.log
( time timeout 900 python bench.py --steps 2 --warmup 3 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/a_bench_reference.json 2> gpurun_out/a_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_bench_n1m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-fmm --no-parity > gpurun_out/a_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uj_direct_f64 -s 1 -c 1 -f -o gpurun_out/r02a_k1_n1m \
    python tools/prof_uj.py 1000000 gaussianerf 2 > gpurun_out/a_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:estr_direct_f64 -s 1 -c 1 -f -o gpurun_out/r02a_k2_n1m \
    python tools/prof_uj.py 1000000 gaussianerf 2 sfs > gpurun_out/a_ncu_k2.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:dfma_peak -c 8 -f -o gpurun_out/r02a_dfma_peak \
    python tools/prof_peak.py > gpurun_out/a_ncu_peak.log 2>&1
tail -3 gpurun_out/a_smoke.log; tail -15 gpurun_out/a_tests.log; cut -c1-1500 gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err; cut -c1-600 gpurun_out/a_bench_reference.json; cat gpurun_out/a_ncu_peak.log | tail -5
