#!/bin/bash
# geometry sweep for the c2 workload: RANKS P BLOCK FILTERS(0 = max concurrent)
run() { echo "== ranks=$1 P=$2 block=$3 filters=$4"; BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 2 1000 2>&1 | grep -E "filters|run ms" | tail -2; }
run 16 4 1024 0
run 16 8 512 0
run 8 8 1024 0
run 8 8 1024 8
run 16 4 1024 14
