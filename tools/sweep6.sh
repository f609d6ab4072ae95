#!/bin/bash
run() { echo "== ranks=$1 P=$2 block=$3 filters=$4"; BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 2 1000 2>&1 | grep -E "run ms" | tail -1; }
run 32 8 256 18
run 64 4 256 18
run 64 4 256 9
run 128 2 256 4
run 16 8 512 18
run 32 4 512 18
