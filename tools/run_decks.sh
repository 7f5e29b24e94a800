#!/bin/bash
# Times rumdeed_b200_run on the example decks of tools/decks/ (copied to a scratch dir; outputs stay there).
# usage: tools/run_decks.sh [steps for the planar decks] [steps for the tip deck]
SP=${1:-2000}; ST=${2:-5000}
EXE=${RB2_RUN_EXE:-$(dirname "$0")/../rumdeed_b200/rumdeed_b200_run}
run() { # name deck steps extra-namelist-line
  d=$(mktemp -d); cp $(dirname "$0")/decks/$2/* $d/
  if [ -n "$4" ]; then sed -i "s|^/|  $4\n/|" $d/input; fi
  s=$(date +%s.%N)
  timeout ${DECK_TIMEOUT:-300} $DECK_PREFIX $EXE $d 20261017 $3 1000000 > $d/log 2>&1; rc=$?
  e=$(date +%s.%N)
  python - "$1" "$d" "$3" "$s" "$e" "$rc" <<'PY'
import sys, numpy as np
name, d, steps, s, e, rc = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6])
log = open(d + "/log").read().strip().splitlines()
try:
    ramo = np.loadtxt(d + "/out/ramo_current.dt")
    I = ramo[int(0.75 * len(ramo)):, 2].mean(); nel = int(ramo[-1, 5])
except Exception as ex:
    I, nel = float("nan"), -1
phase = [l for l in log if "wall clock per phase" in l]
loop = float("nan")
try:  # time inside the main loop (the process start -- CUDA context, first touch of the libraries -- is 2-4 s on a fresh box)
    import re
    m = re.search(r"emission ([0-9.]+)\s+MD step ([0-9.]+)\s+removal ([0-9.]+)\s+writers ([0-9.]+)", phase[0])
    loop = sum(float(x) for x in m.groups())
    c = [l for l in log if "collisions:" in l]
    if c:
        loop += float(re.search(r"wall clock ([0-9.]+) s", c[0]).group(1))
except Exception:
    pass
print(f"{name:28s} rc={rc} steps={steps} wall={e-s:7.2f}s  steps/s={steps/(e-s):8.1f} (main loop only: {steps/loop:8.1f})  I(last quarter)={I:.4e} A  nrElec(end)={nel}")
print("   ", phase[0] if phase else log[-2:])
for l in log:
    if "collisions:" in l or "emission split" in l or "work-unit lists" in l: print("   ", l)
PY
  rm -rf $d
}
WHICH=${3:-all}
if [ "$WHICH" = all ] || [ "$WHICH" = batch ]; then run "GPU-Planar-FE MH_BATCH"   gpu_planar_fe $SP ""; fi
if [ "$WHICH" = all ] || [ "$WHICH" = device ]; then run "GPU-Planar-FE MH_DEVICE"  gpu_planar_fe $SP "MH_DEVICE = .True.,"; fi
if [ "$WHICH" = all ] || [ "$WHICH" = device ]; then run "Planar-FE 4.7eV MH_DEVICE" planar_fe_4p7 $SP "MH_DEVICE = .True.,"; fi
if [ "$WHICH" = all ] || [ "$WHICH" = tip ]; then run "Tip-FE MH_BATCH"          tip_fe $ST "MH_BATCH = .True.,"; fi
if [ "$WHICH" = all ] || [ "$WHICH" = tip ]; then run "Tip-FE MH_DEVICE"         tip_fe $ST "MH_DEVICE = .True.,"; fi
if [ "$WHICH" = all ] || [ "$WHICH" = ion ]; then run "Ion (collisions) MH_DEVICE" ion $SP "MH_DEVICE = .True.,"; fi
if [ "$WHICH" = all ] || [ "$WHICH" = serial ]; then run "Planar-FE 4.7eV as shipped (serial chains, device)" planar_fe_4p7 300 ""; fi
if [ "$WHICH" = all ] || [ "$WHICH" = serial ]; then run "Planar-FE 4.7eV MH_HOST (serial chains, host loop)" planar_fe_4p7 40 "MH_HOST = .True.,"; fi
if [ "$WHICH" = all ] || [ "$WHICH" = photo ]; then run "Photo (500 nm, 300 steps)"  photo 300 ""; fi
if [ "$WHICH" = all ] || [ "$WHICH" = tfe ]; then run "Checkerboard-TFE"          checkerboard_tfe $SP ""; fi
if [ "$WHICH" = all ] || [ "$WHICH" = tfe ]; then run "Checkerboard-TFE MH_DEVICE" checkerboard_tfe $SP "MH_DEVICE = .True.,"; fi
# (the serial default sampler, MH_BATCH = .False., is a host loop of one M = 1 field call per jump: latency bound at
#  ~40 us per call, 2.2 steps/s on the 4.7 eV deck; the reference documents MH_BATCH for GPU builds)
