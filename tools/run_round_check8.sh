#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mh_ -c 2 -o gpurun_out/ncu_mh_small_5000_107 -f python tools/prof_mh_small.py 5000 107 > gpurun_out/ncu_mh8.log 2>&1; tail -5 gpurun_out/ncu_mh8.log
