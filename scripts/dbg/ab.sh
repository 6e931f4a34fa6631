#!/bin/bash
# A/B of variant builds of libgf2_b200.so: scripts/dbg/ab.sh v1 v2 ...   (names under scripts/dbg/variants/, without .so)
for v in "$@"; do
  GF2_LIB=scripts/dbg/variants/$v.so timeout 120 python scripts/dbg/lin_perf.py 4096 2 2>&1 | grep -v "^$" | tail -22
done
