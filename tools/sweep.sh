#!/bin/bash
# geometry sweep for the c2 workload: RANKS P BLOCK FILTERS
run() { echo "== ranks=$1 P=$2 block=$3 filters=$4"; BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 3 1000 2>&1 | tail -2; }
run 16 4 1024 1
run 16 4 1024 8
run 16 8 512 1
run 16 8 512 8
run 16 8 512 16
run 8 8 1024 16
run 8 8 1024 1
