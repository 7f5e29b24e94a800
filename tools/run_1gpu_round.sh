#!/bin/bash
# Single-GPU round check: all GPU tests, headline bench, launch list of the bench command, DRAM traffic at N = 1e6.
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1_tests.log; tail -2 gpurun_out/r1_tests.log
python bench.py > gpurun_out/r1_bench_1e6.json 2> gpurun_out/r1_bench.err
python bench.py --particles 100000 --steps 5 > gpurun_out/r1_bench_1e5.json 2>> gpurun_out/r1_bench.err
python bench.py --particles 10000 --steps 20 > gpurun_out/r1_bench_1e4.json 2>> gpurun_out/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_bench1e5.csv python bench.py --particles 100000 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1_ncu_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dram_r01_n1e6_t2.csv python tools/prof_step.py --n 1000000 --steps 1 > gpurun_out/r1_ncu_dram.log 2>&1
tail -c 600 gpurun_out/r1_bench_1e6.json
