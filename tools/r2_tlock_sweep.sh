# usage (GPU box): bash tools/r2_tlock_sweep.sh -- run kernel with the ring hand-off vs with T-row locks (MFB200_KERNEL=tlock)
# on one rank's share of C3 per sub-step on 8 / 4 / 2 GPUs (timed alone on one GPU) and on the named shapes
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); printf "%s %s", g, b} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
t run 60000,2225,1562500,128 MFB200_KERNEL=run
for c in 78 111 148; do for s1 in 1 2 4; do
  t tlock 60000,2225,1562500,128 MFB200_KERNEL=tlock MFB200_RING_CTAS=$c MFB200_RING_S1=$s1
done; done
t tlock8w 60000,2225,1562500,128 MFB200_KERNEL=tlock MFB200_RING_WARPS=8
for shape in 120000,4450,6250000,128 240000,8900,25000000,128 c3 c2 c1; do
  t run $shape MFB200_KERNEL=run
  for s1 in 1 2; do t tlock $shape MFB200_KERNEL=tlock MFB200_RING_S1=$s1; done
done
MFB200_STATS=1 MFB200_KERNEL=tlock python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | grep -E "stats|epoch 3"
MFB200_STATS=1 MFB200_KERNEL=tlock python tools/prof_ring.py c3 4 2>&1 | grep -E "stats|epoch 3"
