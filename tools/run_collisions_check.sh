#!/bin/bash
# GPU check of the collision step: its parity tests, the recombination / ionisation micro-benchmark beside the CPU
# oracle, the Ion deck, an ncu capture of the ion x electron sweep, then the full GPU suite (regression).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collisions.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/coll_tests.log; tail -3 gpurun_out/coll_tests.log
timeout 300 python tools/bench_recomb.py 10000 100000 1000000 --cpu > gpurun_out/bench_recomb.log 2>&1; cat gpurun_out/bench_recomb.log
DECK_TIMEOUT=200 timeout 240 tools/run_decks.sh 2000 5000 ion > gpurun_out/deck_ion.log 2>&1; cat gpurun_out/deck_ion.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_recomb_sweep -s 2 -c 1 -o gpurun_out/ncu_recomb_sweep -f python tools/bench_recomb.py 1000000 > gpurun_out/ncu_recomb.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_recomb.csv python tools/bench_recomb.py 100000 > gpurun_out/ncu_recomb_launches.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/all_tests.log; tail -3 gpurun_out/all_tests.log
