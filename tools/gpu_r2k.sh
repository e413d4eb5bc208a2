#!/bin/bash
# Round-2 GPU pass K (8 GPUs): per-rank device time of every FMM section + clocks/power of every GPU during the run.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 100 > gpurun_out/k_smi.csv 2>/dev/null &
SMI=$!
( time timeout 240 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --uj fmm --particles 5000000 --steps 5 --warmup 5 --let-timing --no-parity ) > gpurun_out/k_bench_fmm_5m_8gpu.json 2> gpurun_out/k_bench_fmm_5m_8gpu.err
kill $SMI
