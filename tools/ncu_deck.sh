d=$(mktemp -d); cp tools/decks/gpu_planar_fe/* $d/; sed -i "s|^/|  MH_DEVICE = .True.,\n/|" $d/input
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 24000 --launch-count 600 --csv --log-file gpurun_out/launches_deck.csv rumdeed_b200/rumdeed_b200_run $d 20261017 1500 1000000 > gpurun_out/ncu_deck.log 2>&1
tail -3 gpurun_out/ncu_deck.log
