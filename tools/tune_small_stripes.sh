#!/bin/bash
# usage (GPU box): tools/tune_small_stripes.sh <tag>
# One rank's share of C3 per sub-step when the item stripes rotate over G GPUs (G = 2, 4, 8), trained on ONE GPU:
# the launch shape of the multi-GPU path without the NCCL hand-off.  Sweeps lanes per rating and minimum cell size;
# a second run of the default shape with MFB200_STATS=1 prints the scheduling counters (slower kernel variant).
tag=$1
out=gpurun_out/${tag}_small.log
: > $out
for shape in 240000,8900,25000000,128 120000,4450,6250000,128 60000,2225,1562500,128; do
  for L in 8 16 32; do
    for mc in 4 1; do
      echo "== shape $shape L=$L min_cell=$mc" >> $out
      MFB200_GROUP_LANES=$L MFB200_MIN_CELL=$mc timeout 120 python tools/prof_ring.py $shape 4 2>&1 | tail -3 >> $out
    done
  done
  echo "== shape $shape default, counters" >> $out
  MFB200_STATS=1 timeout 120 python tools/prof_ring.py $shape 4 2>&1 | grep -E "stats|epoch 3" >> $out
done
grep -E "==|epoch 3|grid_ctas|stats" $out | cut -c1-420
