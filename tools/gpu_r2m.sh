#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_branch.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum \
   --clock-control none -k regex:fmm_leaf_uj -s 55 -c 24 --csv --log-file gpurun_out/m_leaf_uj_ranks.csv python tools/let_balance_probe.py 5000000 8 rings 11 > gpurun_out/m_probe.log 2>&1
tail -3 gpurun_out/m_probe.log
