#!/bin/bash
# Launch list of a deck's late steps: tools/ncu_deck.sh [deck] [steps] [launches to skip]
DECK=${1:-gpu_planar_fe}; STEPS=${2:-1500}; SKIP=${3:-24000}
d=$(mktemp -d); cp tools/decks/$DECK/* $d/; sed -i "s|^/|  MH_DEVICE = .True.,\n/|" $d/input
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP --launch-count 400 --csv --log-file gpurun_out/launches_deck_$DECK.csv rumdeed_b200/rumdeed_b200_run $d 20261017 $STEPS 1000000 > gpurun_out/ncu_deck.log 2>&1
tail -2 gpurun_out/ncu_deck.log
python - "$DECK" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_deck_%s.csv" % sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
d = collections.OrderedDict()
for r in rows:
    k = r[4][:56]
    d.setdefault(k, []).append(float(r[-1]))
for k, v in d.items():
    print(f"{k:58s} n={len(v):4d} avg {sum(v)/len(v)/1e3:9.2f} us total {sum(v)/1e6:8.3f} ms")
PY
