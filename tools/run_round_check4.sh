#!/bin/bash
mkdir -p gpurun_out
DECK_TIMEOUT=200 timeout 400 tools/run_decks.sh 2000 5000 device > gpurun_out/deck_device4.log 2>&1; cat gpurun_out/deck_device4.log
DECK_TIMEOUT=200 timeout 240 tools/run_decks.sh 2000 5000 ion > gpurun_out/deck_ion4.log 2>&1; cat gpurun_out/deck_ion4.log
