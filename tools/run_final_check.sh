#!/bin/bash
# Round-end check on one B200: smoke, the whole GPU suite, the bench line (+ reference arm), launch list and full ncu
# capture of the dominant kernel, all example decks.  ~12 minutes.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -25 | tee gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench_1e6.json 2> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench_1e6.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench_ref.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_bench1e6.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/final_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_sym -s 1 -c 1 -f -o gpurun_out/ncu_pairsym_r02_n1e5 python tools/prof_step.py --n 100000 --steps 2 > gpurun_out/final_ncu1.log 2>&1
bash tools/run_decks.sh 2000 5000 > gpurun_out/final_decks.log 2>&1; cat gpurun_out/final_decks.log
