#!/bin/bash
# A/B of the GPU-Planar-FE deck's MD-step phase between library builds (same box)
for v in 0b794ef 4dcfc8d head 0b794ef head; do
  echo "== $v"
  if [ $v = head ]; then unset RB2_RUN_EXE; else export RB2_RUN_EXE=$PWD/tools/variants/$v/rumdeed_b200_run; fi
  bash tools/run_decks.sh 2000 5000 device 2>&1 | grep -A1 "GPU-Planar-FE" 
done
