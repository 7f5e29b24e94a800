#!/bin/bash
# Round-2 check #3: the tests touched since #2, N = 1e4 step with / without the step graph, op mix of both pair kernels.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_c_boundary.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2c3_tests.log; tail -3 gpurun_out/r2c3_tests.log
python bench.py --particles 10000 --steps 30 --no-cpu --no-sweep > gpurun_out/r2c3_bench_1e4.json 2> gpurun_out/r2c3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3_bench_1e4.json'))
print('1e4 graph: ms/step', d['ms_per_step'], 'median', sorted(d['ms_steps_rank0'])[15], 'kernel', d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
PY
RB2_NO_GRAPH=1 python bench.py --particles 10000 --steps 30 --no-cpu --no-sweep > gpurun_out/r2c3_bench_1e4_nograph.json 2>> gpurun_out/r2c3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3_bench_1e4_nograph.json'))
print('1e4 plain: ms/step', d['ms_per_step'], 'median', sorted(d['ms_steps_rank0'])[15], 'kernel', d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
PY
python tools/variant_bench.py 10000 > gpurun_out/r2c3_variants_1e4.log 2>&1; cat gpurun_out/r2c3_variants_1e4.log
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum,smsp__thread_inst_executed_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:"k_pair" --csv --log-file gpurun_out/opmix_r02_n1e5_sym.csv python tools/prof_step.py --n 100000 --steps 1 > gpurun_out/r2c3_ncu1.log 2>&1
RB2_PAIR_MODE=1 ncu --metrics $M --clock-control none -k regex:"k_pair" --csv --log-file gpurun_out/opmix_r02_n1e5_gather.csv python tools/prof_step.py --n 100000 --steps 1 > gpurun_out/r2c3_ncu2.log 2>&1
tail -3 gpurun_out/opmix_r02_n1e5_sym.csv; tail -3 gpurun_out/opmix_r02_n1e5_gather.csv
