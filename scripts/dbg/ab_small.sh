#!/bin/bash
for v in "$@"; do for B in 1 8 148; do
  GF2_LIB=scripts/dbg/variants/$v.so timeout 120 python scripts/dbg/lin_perf.py $B 5 2>&1 | tail -1 | cut -c1-200
done; done
