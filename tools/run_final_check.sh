#!/bin/bash
# Round-end check on one GPU: smoke, the whole GPU suite, the headline bench lines and the deck runs.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench_1e6.json 2> gpurun_out/final_bench.err; tail -c 1500 gpurun_out/final_bench_1e6.json
timeout 300 python bench.py --particles 100000 --steps 5 > gpurun_out/final_bench_1e5.json 2>> gpurun_out/final_bench.err
timeout 300 python bench.py --particles 10000 --steps 20 > gpurun_out/final_bench_1e4.json 2>> gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench_ref.json
DECK_TIMEOUT=200 timeout 700 tools/run_decks.sh 2000 5000 all > gpurun_out/final_decks.log 2>&1; cat gpurun_out/final_decks.log
