#!/bin/bash
# A/B timing of Legendre kernel variants: the in-tree library, then variants/*.so
python scripts/time_leg.py
for f in variants/*.so; do [ -f $f ] && PLK_LIB_PATH=$f python scripts/time_leg.py; done
