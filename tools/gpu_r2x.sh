#!/bin/bash
# Round-2 GPU pass X (2 GPUs): per-step probe of the all-gather LET mode (the 917 ms/step line of pass W).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29521 tools/let_step_probe.py 5000000 let 8 1 > gpurun_out/x_probe_let.log 2> gpurun_out/x_probe_let.err
grep -c . gpurun_out/x_probe_let.log; cut -c1-420 gpurun_out/x_probe_let.log; tail -3 gpurun_out/x_probe_let.err
