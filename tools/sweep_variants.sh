#!/bin/bash
# times each build variant on the C2 batch, C2 single-filter and C4 workloads
for so in build_variants/*.so; do
  echo "== $so"
  for nf in ${FILTERS:-18 1}; do
    echo -n "filters=$nf: "; BSSM_B200_LIB=$PWD/$so python tools/profile_run.py $nf 3 1000 2>&1 | grep -E "run ms" | tail -1
  done
  echo -n "c4: "; BSSM_B200_LIB=$PWD/$so python tools/profile_c4.py 2>&1 | tail -1
done
