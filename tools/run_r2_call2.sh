#!/bin/bash
# Round-2 check #2: full GPU suite, headline bench line with sweep + field batches, ncu captures of k_pair_sym, decks.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c2_tests.log; tail -3 gpurun_out/r2c2_tests.log
python bench.py > gpurun_out/r2c2_bench_1e6.json 2> gpurun_out/r2c2_bench.err; tail -c 1500 gpurun_out/r2c2_bench_1e6.json; tail -3 gpurun_out/r2c2_bench.err
ncu --set full --clock-control none --import-source on -k regex:k_pair_sym -s 2 -c 1 -f -o gpurun_out/ncu_pairsym_r02_n1e5 python tools/prof_step.py --n 100000 --steps 2 > gpurun_out/r2c2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_sym -s 0 -c 1 -f -o gpurun_out/ncu_pairsym_r02_n1e6 python tools/prof_step.py --n 1000000 --steps 1 > gpurun_out/r2c2_ncu2.log 2>&1
bash tools/run_decks.sh 2000 5000 > gpurun_out/r2c2_decks.log 2>&1; cat gpurun_out/r2c2_decks.log
