#!/bin/bash
mkdir -p gpurun_out
for n in 10000 100000 1000000; do python tools/variant_bench.py $n > gpurun_out/r2c4_variants_$n.log 2>&1; cat gpurun_out/r2c4_variants_$n.log; done
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "close_pairs or symmetric" 2>&1 | tail -3
