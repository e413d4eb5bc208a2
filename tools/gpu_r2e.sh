#!/bin/bash
# Round-2 GPU pass E (1 GPU): launch list of one UJ_fmm step through the LET path (world = 1 runs every phase) at 2.5M particles
# (= an 2-GPU rank's share of 5M), and the multi / let / simloop tests on one GPU (all shards on cuda:0).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02e_launches_fmm_let_2p5m.csv \
    python bench.py --uj fmm --particles 2500000 --steps 1 --warmup 1 --no-parity > gpurun_out/e_ncu.log 2>&1
( time timeout 300 python bench.py --uj fmm --particles 625000 --steps 3 --warmup 3 --no-parity --let-timing ) > gpurun_out/e_bench_fmm_625k.json 2> gpurun_out/e_bench_fmm_625k.err
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_let.py -q ) > gpurun_out/e_tests.log 2>&1
tail -3 gpurun_out/e_tests.log; cut -c1-1500 gpurun_out/e_bench_fmm_625k.json
