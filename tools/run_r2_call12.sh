#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_emission.py -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -30 | tee gpurun_out/r2c12_tests.log
bash tools/run_decks.sh 2000 5000 serial 2>&1 | tee gpurun_out/r2c12_decks.log
