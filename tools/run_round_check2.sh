#!/bin/bash
# Second GPU check of this session: collision + emission tests, the single-barrier sampler against the cooperative one,
# recombination sweep with two ions per thread, the decks, then the rest of the GPU suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_collisions.py tests/test_emission.py -m gpu -q 2>&1 | tail -40 > gpurun_out/c2_tests.log; tail -4 gpurun_out/c2_tests.log
timeout 300 python tools/bench_mh_small.py > gpurun_out/bench_mh_small.log 2>&1; cat gpurun_out/bench_mh_small.log
timeout 200 python tools/bench_recomb.py 10000 100000 1000000 > gpurun_out/bench_recomb2.log 2>&1; cat gpurun_out/bench_recomb2.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device2.log 2>&1; cat gpurun_out/deck_device2.log
DECK_TIMEOUT=200 timeout 240 tools/run_decks.sh 2000 5000 ion > gpurun_out/deck_ion2.log 2>&1; cat gpurun_out/deck_ion2.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/c2_rest.log; tail -3 gpurun_out/c2_rest.log
