#!/bin/bash
# the four numbers tracked while tuning: C2 batch, C2 single filter, C3 (growth), C5 (sde)
echo "== c2 batch 18"; python tools/profile_run.py 18 3 1000 2>&1 | grep -E "run ms" | tail -1
echo "== c2 single"; python tools/profile_run.py 1 3 1000 2>&1 | grep -E "run ms" | tail -1
python tools/profile_c3c5.py 2>&1 | tail -6
