#!/bin/bash
for P in 2 4 8; do echo "== C4 P=$P"; BSSM_B200_P=$P python tools/profile_c4.py 2>&1 | tail -2; done
