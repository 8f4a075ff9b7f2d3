#!/bin/bash
# usage (GPU box): tools/sweep_small.sh <lib relative to the package> <shape> -- sweeps CTAs x slack x warps for one launch shape
PKG="$(cd "$(dirname "$0")/.." && pwd)/question-recommendation-system_b200"
lib=$1; shape=$2
for w in 16 8; do for s1 in 1 2; do for c in 32 48 64 78 110 148; do
  r=$(MFB200_LIB="$PKG/$lib" MFB200_RING_WARPS=$w MFB200_RING_S1=$s1 MFB200_RING_CTAS=$c timeout 100 python "$(dirname "$0")/prof_ring.py" $shape 5 2>&1 | grep -E "epoch 4" | awk '{print $3}')
  echo "$shape warps=$w S1=$s1 ctas=$c ms=$r"
done; done; done
