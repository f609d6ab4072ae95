#!/bin/bash
run() { echo "== ranks=$1 P=$2 block=$3 filters=$4 sync=$5"; BSSM_B200_SYNC=$5 BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 2 1000 2>&1 | grep -E "filters|run ms" | tail -2; }
echo "== defaults"; python tools/profile_run.py 1 2 1000 2>&1 | grep "run ms" | tail -1
python tools/profile_run.py 18 2 1000 2>&1 | grep "run ms" | tail -1
run 128 2 256 1 group
run 64 4 256 1 group
run 64 2 512 1 group
run 32 2 1024 1 group
python tools/profile_c3c5.py c3 2 | tail -1
python tools/profile_c3c5.py c5 2 | tail -1
python tools/profile_c4.py 15000 2 | tail -1
