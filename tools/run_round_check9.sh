#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_emission.py -m gpu -q -x -k "single_barrier or device_resident" 2>&1 | tail -30 > gpurun_out/c9_tests.log; tail -4 gpurun_out/c9_tests.log
MH_M=10,64,107,128 timeout 300 python tools/bench_mh_small.py 0 1000 5000 9000 20000 > gpurun_out/bench_mh9.log 2>&1; cut -c1-100 gpurun_out/bench_mh9.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device9.log 2>&1; cat gpurun_out/deck_device9.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mh_ -c 2 -o gpurun_out/ncu_mh_small_5000_107 -f python tools/prof_mh_small.py 5000 107 > gpurun_out/ncu_mh9.log 2>&1; tail -5 gpurun_out/ncu_mh9.log
