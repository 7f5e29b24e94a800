#!/bin/bash
mkdir -p gpurun_out
for v in base closeoff; do echo "== $v"; RB2_LIB_PATH=tools/variants/$v/librumdeed_b200.so MH_M=107,324 python tools/bench_mh_small.py 1300 9000 2>&1 | cut -c1-120; done | tee gpurun_out/r2c5_mh.log
bash tools/run_decks.sh 2000 5000 device 2>&1 | tee gpurun_out/r2c5_decks.log
