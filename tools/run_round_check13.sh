#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_emission.py -m gpu -q -x -k "device_resident_tip" 2>&1 | tail -30 > gpurun_out/c13_tip.log; tail -8 gpurun_out/c13_tip.log
