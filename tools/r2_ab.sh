#!/bin/bash
# usage (GPU box): tools/r2_ab.sh <tag> [kernels] -- band kernel vs run kernel on the named shapes and on one rank's
# share of C3 per sub-step when the item stripes rotate over 2 / 4 / 8 GPUs (trained alone on one GPU), plus counters.
tag=$1
kernels=${2:-"band run"}
out=gpurun_out/${tag}_ab.log
: > $out
for shape in c3 c2 c1 240000,8900,25000000,128 120000,4450,6250000,128 60000,2225,1562500,128; do
  for kern in $kernels; do
    echo "== shape $shape kernel=$kern" >> $out
    MFB200_KERNEL=$kern timeout 300 python tools/prof_ring.py $shape 6 2>&1 | grep -E "epoch [45]|grid_ctas" | cut -c1-200 >> $out
  done
done
for shape in c3 60000,2225,1562500,128; do
  echo "== shape $shape kernel=run counters" >> $out
  MFB200_STATS=1 MFB200_KERNEL=run timeout 300 python tools/prof_ring.py $shape 4 2>&1 | grep -E "stats|epoch 3" | tail -2 >> $out
done
cat $out | cut -c1-400
