#!/bin/bash
run() { echo "== wide=$6 ranks=$1 P=$2 block=$3 filters=$4"; BSSM_B200_WIDE=$6 BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 2 1000 2>&1 | grep -E "filters|run ms" | tail -2; }
run 32 8 256 0 group 1
run 32 8 256 0 group 0
run 64 4 256 0 group 1
run 64 4 256 0 group 0
BSSM_B200_WIDE=1 BSSM_B200_P=8 BSSM_B200_BLOCK=32 python tools/profile_c4.py 15000 2 | tail -1
BSSM_B200_WIDE=1 BSSM_B200_P=4 BSSM_B200_BLOCK=64 python tools/profile_c4.py 15000 2 | tail -1
BSSM_B200_P=4 BSSM_B200_BLOCK=64 python tools/profile_c4.py 15000 2 | tail -1
