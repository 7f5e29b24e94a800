#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_emission.py -m gpu -q -x -k "single_barrier or device_resident or planar_system" 2>&1 | tail -30 > gpurun_out/c16_tests.log; tail -4 gpurun_out/c16_tests.log
MH_M=10,107,324 timeout 200 python tools/bench_mh_small.py 0 1000 5000 9000 > gpurun_out/bench_mh16.log 2>&1; cut -c1-100 gpurun_out/bench_mh16.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device16.log 2>&1; grep -h "steps/s\|emission split" gpurun_out/deck_device16.log
