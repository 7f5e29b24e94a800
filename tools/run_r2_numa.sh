#!/bin/bash
# Is the bimodal MD-step phase of the decks a matter of which CPUs the process runs on?
nvidia-smi topo -m 2>&1 | head -8
lscpu | grep -i "numa\|^CPU(s)\|Model name\|Thread"
nproc; taskset -p $$
LOCAL=$(nvidia-smi topo -m | awk '/^GPU0/{for(i=1;i<=NF;i++) if ($i ~ /^[0-9]+-[0-9]+(,[0-9]+-[0-9]+)*$/) {print $i; exit}}')
echo "GPU0 local cpus: $LOCAL"
run() { bash tools/run_decks.sh 2000 5000 device 2>&1 | grep -A1 "GPU-Planar-FE" | tail -1 | sed 's/.*wall clock per phase/   phase/'; }
for k in 1 2 3; do echo "== free"; run; done
if [ -n "$LOCAL" ]; then for k in 1 2 3; do echo "== taskset $LOCAL"; DECK_PREFIX="taskset -c $LOCAL" run; done; fi
