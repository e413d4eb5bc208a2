#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python tools/let_balance_probe.py 5000000 8 rings 11 ) > gpurun_out/l_probe_rings_evolved.log 2>&1
grep -v "^rank" gpurun_out/l_probe_rings_evolved.log | tail -20; grep "^rank" gpurun_out/l_probe_rings_evolved.log | tail -8
