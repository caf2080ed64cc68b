#!/bin/bash
# A/B of experimental builds under variants/ against the in-tree library: kernel timings + a correctness subset.
#   variants are built by hand, e.g.  nvcc <flags of _build.py> -DPLK_ANA_DEFER=1 -o variants/defer.so plk_api.cu
python scripts/time_leg.py
python scripts/time_ring.py
for f in variants/*.so; do
  [ -f $f ] || continue
  PLK_LIB_PATH=$f python scripts/time_leg.py
  PLK_LIB_PATH=$f python scripts/time_ring.py
  PLK_LIB_PATH=$f timeout 120 python -m pytest tests/test_sht_gpu.py -q -x -m gpu -k "analysis_matches_oracle or adjointness_full_size or mid_size" 2>&1 | tail -2
done
