#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collisions.py "tests/test_emission.py::test_device_sampler_single_barrier_kernel" -m gpu -q 2>&1 | tail -30 > gpurun_out/c3_tests.log; tail -4 gpurun_out/c3_tests.log
timeout 300 python tools/bench_mh_small.py 0 1000 5000 10000 19000 40000 > gpurun_out/bench_mh_small3.log 2>&1; cat gpurun_out/bench_mh_small3.log
RB2_MH_SMALL=1 DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device3_small1.log 2>&1; cat gpurun_out/deck_device3_small1.log
RB2_MH_SMALL=0 DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device3_small0.log 2>&1; cat gpurun_out/deck_device3_small0.log
bash tools/ncu_deck.sh
