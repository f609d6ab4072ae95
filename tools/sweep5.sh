#!/bin/bash
python tools/profile_run.py 1 2 1000 2>&1 | grep "run ms" | tail -1
python tools/profile_run.py 18 2 1000 2>&1 | grep "run ms" | tail -1
python tools/profile_c4.py 15000 2 | tail -1
python tools/profile_c3c5.py c3 2 | tail -1
python tools/profile_c3c5.py c5 2 | tail -1
