#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_emission.py tests/test_gpu_parity.py -m gpu -q -k "serial or planar_system or thermo or refuses or checkerboard" 2>&1 | grep -v "^    \|^$" | tail -12
bash tools/run_decks.sh 2000 5000 serial 2>&1 | grep -v MH_HOST -A2 | head -4
