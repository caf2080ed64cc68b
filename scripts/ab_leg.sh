#!/bin/bash
# NR sweep of the in-tree library, then variants/*.so
python scripts/time_leg.py
PLK_NR_SYN0=2 PLK_NR_SYNS=2 PLK_NR_ANA0=2 PLK_NR_ANAS=1 python scripts/time_leg.py
for f in variants/*.so; do [ -f $f ] && PLK_LIB_PATH=$f python scripts/time_leg.py; done
