#!/bin/bash
# One GPU-box validation pass (run as: gpurun --timeout 1800 -- 'bash tools/gpu_validate.sh'): smoke, the GPU test suite,
# both bench arms, the workload sweep, and a launch list of the FMM path.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/v_smoke.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/v_tests.log 2>&1
echo "tests exit: $?" >> gpurun_out/v_tests.log
( time timeout 900 python bench.py ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/v_bench_reference.json 2> gpurun_out/v_bench_reference.err
rm -f gpurun_out/sweep.jsonl
( time timeout 900 python tools/sweep.py --cases wing,rotor,vahana,random --max-n ${SWEEP_MAX_N:-50000000} ) > gpurun_out/v_sweep.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches_fmm_n1m.csv python tools/prof_fmm.py 1000000 > gpurun_out/v_ncu.log 2>&1
tail -3 gpurun_out/v_smoke.log; tail -4 gpurun_out/v_tests.log; tail -2 gpurun_out/v_sweep.log | cut -c1-300; cut -c1-400 gpurun_out/v_bench.json
