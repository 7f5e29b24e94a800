#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_emission.py -m gpu -q -k "sampler or tip or planar_system or thermo" 2>&1 | tail -5
bash tools/run_decks.sh 2000 5000 tip 2>&1 | grep -v BATCH -A3 | tee gpurun_out/r2c7_decks.log
bash tools/run_decks.sh 2000 5000 device 2>&1 | tee -a gpurun_out/r2c7_decks.log
