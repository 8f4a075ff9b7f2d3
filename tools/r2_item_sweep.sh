# usage (GPU box): bash tools/r2_item_sweep.sh -- the item kernel (MFB200_KERNEL=item; 8 or 32 lanes per S row) against the default
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 300 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.cta_warps.: [0-9]+/); w=substr($0,RSTART,RLENGTH); match($0,/.kernel.: [0-9]+/); k=substr($0,RSTART,RLENGTH); printf "%s %s %s", g, w, k} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
for shape in 60000,2225,1562500,128 120000,4450,6250000,128 240000,8900,25000000,128 c3 c2 c1 c4; do
  t default $shape
  t item32 $shape MFB200_KERNEL=item
  t item8 $shape MFB200_KERNEL=item MFB200_ITEM_LANES=8
done
for w in 8 16; do t item32_w$w 60000,2225,1562500,128 MFB200_KERNEL=item MFB200_RING_WARPS=$w; done
MFB200_STATS=1 MFB200_KERNEL=item python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | grep -E "stats|epoch 3" | tail -2
MFB200_STATS=1 MFB200_KERNEL=item python tools/prof_ring.py c3 4 2>&1 | grep -E "stats|epoch 3" | tail -2
