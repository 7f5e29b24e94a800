#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_emission.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2c6_tests.log; tail -5 gpurun_out/r2c6_tests.log
bash tools/run_decks.sh 2000 5000 tip 2>&1 | tee gpurun_out/r2c6_decks.log
