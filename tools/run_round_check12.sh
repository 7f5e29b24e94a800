#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_emission.py -m gpu -q -x -k "tip" 2>&1 | tail -30 > gpurun_out/c12_tip.log; tail -5 gpurun_out/c12_tip.log
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 tip > gpurun_out/deck_tip12.log 2>&1; cat gpurun_out/deck_tip12.log
