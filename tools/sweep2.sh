#!/bin/bash
run() { echo "== ranks=$1 P=$2 block=$3 filters=$4"; BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 2 1000 2>&1 | grep -E "run ms" | tail -1; }
run 32 2 1024 1
run 64 2 512 1
run 128 2 256 1
run 64 4 256 1
run 128 4 128 1
run 256 2 128 1
