#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_emission.py -m gpu -q -x -k "single_barrier" 2>&1 | tail -30 > gpurun_out/c14_tests.log; tail -6 gpurun_out/c14_tests.log
MH_M=107,200,324,512 timeout 200 python tools/bench_mh_small.py 1000 5000 9000 > gpurun_out/bench_mh14.log 2>&1; cut -c1-100 gpurun_out/bench_mh14.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device14.log 2>&1; cat gpurun_out/deck_device14.log
