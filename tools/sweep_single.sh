#!/bin/bash
# single C2 filter (N = 65536, T = 1000): kernel ms for forced geometries (default: 128 CTAs x 256 threads, P = 2, group sync)
for spec in "16 4 cluster" "8 8 cluster" "16 4 group" "32 2 group" "64 2 group" "128 2 group" "256 2 group"; do
  set -- $spec
  out=$(BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_SYNC=$3 python tools/profile_run.py 1 3 1000 2>&1 | grep "run ms" | tail -1 | awk '{print $3}')
  echo "ranks=$1 P=$2 sync=$3: $out ms"
done
