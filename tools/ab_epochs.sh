#!/bin/bash
# A/B of kernel variants on one GPU: epoch times of the named workloads with every library given.
# usage: tools/ab_epochs.sh "c3 c2 c1" lib/libmf.so lib_f2/libmf.so ...   (paths relative to the package)
PKG="$(cd "$(dirname "$0")/.." && pwd)/question-recommendation-system_b200"
WLS="$1"; shift
for wl in $WLS; do
  for lib in "$@"; do
    echo "== $wl $lib"
    MFB200_LIB="$PKG/$lib" python "$(dirname "$0")/prof_ring.py" "$wl" 6 2>&1 | grep -E "epoch [2-5]|Error|error" 
  done
done
