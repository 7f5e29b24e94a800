#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_emission.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c7_tests.log; tail -4 gpurun_out/c7_tests.log
MH_M=10,64,107,128 timeout 300 python tools/bench_mh_small.py 0 1000 5000 9000 20000 > gpurun_out/bench_mh7.log 2>&1; cut -c1-100 gpurun_out/bench_mh7.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device7.log 2>&1; cat gpurun_out/deck_device7.log
