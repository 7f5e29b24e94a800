#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_emission.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c5_tests.log; tail -4 gpurun_out/c5_tests.log
MH_M=10,100,324,1000 timeout 300 python tools/bench_mh_small.py 0 1000 10000 40000 > gpurun_out/bench_mh5.log 2>&1; cut -c1-100 gpurun_out/bench_mh5.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device5.log 2>&1; cat gpurun_out/deck_device5.log
RB2_MH_CTAS_PER_SM=1 DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device5_c1.log 2>&1; cat gpurun_out/deck_device5_c1.log
RB2_MH_CTAS_PER_SM=4 DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device5_c4.log 2>&1; cat gpurun_out/deck_device5_c4.log
