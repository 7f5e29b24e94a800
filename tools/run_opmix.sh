#!/bin/bash
# Opcode mix of the two pair kernels at N = 1e5 (the numbers behind bench.py's EXECUTED table), ~1 minute
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum,smsp__thread_inst_executed_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"
TAG=${1:-r02b}
ncu --metrics $M --clock-control none -k regex:"k_pair" --csv --log-file gpurun_out/opmix_${TAG}_n1e5_sym.csv python tools/prof_step.py --n 100000 --steps 1 > gpurun_out/opmix_ncu1.log 2>&1
RB2_PAIR_MODE=1 ncu --metrics $M --clock-control none -k regex:"k_pair" --csv --log-file gpurun_out/opmix_${TAG}_n1e5_gather.csv python tools/prof_step.py --n 100000 --steps 1 > gpurun_out/opmix_ncu2.log 2>&1
tail -8 gpurun_out/opmix_${TAG}_n1e5_sym.csv | cut -d, -f5,13-; tail -8 gpurun_out/opmix_${TAG}_n1e5_gather.csv | cut -d, -f5,13-
