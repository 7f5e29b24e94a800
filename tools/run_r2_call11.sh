#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "symmetric or large_cloud or full_size or close_pairs or graph" 2>&1 | tail -3
for n in 10000 30000 100000 1000000; do RB2_LIB_PATH= python tools/variant_bench.py $n 2>&1 | grep "^base"; done
python tools/sym_unit_sweep.py 1000000 2>&1 | grep "budget=2048 waves=16 kmax=12 gmax=24\|budget=2048 waves=32 kmax=12 gmax=24"
