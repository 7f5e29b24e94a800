#!/bin/bash
# Round-2 check #1: GPU test suite, close-pair flag variants at N = 1e5, deck runs incl. the two new decks.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c1_tests.log; tail -3 gpurun_out/r2c1_tests.log
python tools/variant_bench.py 100000 > gpurun_out/r2c1_variants_1e5.log 2>&1; cat gpurun_out/r2c1_variants_1e5.log
python tools/variant_bench.py 1000000 > gpurun_out/r2c1_variants_1e6.log 2>&1; cat gpurun_out/r2c1_variants_1e6.log
