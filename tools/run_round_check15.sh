#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; tail -1 gpurun_out/final2_smoke.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/final2_tests.log; tail -3 gpurun_out/final2_tests.log
DECK_TIMEOUT=200 timeout 300 tools/run_decks.sh 2000 5000 ion > gpurun_out/final2_deck_ion.log 2>&1; cat gpurun_out/final2_deck_ion.log
