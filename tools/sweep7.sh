#!/bin/bash
run() { echo "== ranks=$1 P=$2 block=$3 filters=$4"; BSSM_B200_RANKS=$1 BSSM_B200_P=$2 BSSM_B200_BLOCK=$3 python tools/profile_run.py $4 3 1000 2>&1 | grep -E "run ms" | tail -1; }
echo "== default 18"; python tools/profile_run.py 18 3 1000 2>&1 | grep -E "run ms" | tail -1
echo "== default 1"; python tools/profile_run.py 1 3 1000 2>&1 | grep -E "run ms" | tail -1
run 64 8 128 18
run 64 8 128 9
run 16 8 512 18
run 64 4 256 9
run 256 2 128 1
run 148 2 224 1
run 64 4 256 1
