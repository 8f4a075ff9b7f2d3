# usage (GPU box): bash tools/r2_tlock_sweep4.sh -- extra tries for a busy S row (MFB200_S_SPIN), lock-ahead depth (MFB200_TL_AHEAD)
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); printf "%s", g} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
for shape in 60000,2225,1562500,128 120000,4450,6250000,128; do
for sp in 0 2 4 8 16; do
  t run $shape MFB200_KERNEL=run MFB200_S_SPIN=$sp
  t run148 $shape MFB200_KERNEL=run MFB200_S_SPIN=$sp MFB200_RING_CTAS=148
  for ah in 1 2; do t tlock $shape MFB200_KERNEL=tlock MFB200_S_SPIN=$sp MFB200_TL_AHEAD=$ah; done
done; done
for shape in 240000,8900,25000000,128 c3 c2 c1; do
  for sp in 0 4; do t run $shape MFB200_KERNEL=run MFB200_S_SPIN=$sp; t tlock $shape MFB200_KERNEL=tlock MFB200_S_SPIN=$sp; done
done
MFB200_STATS=1 MFB200_KERNEL=tlock MFB200_S_SPIN=8 MFB200_TL_AHEAD=1 python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | grep -E "stats|epoch 3" | tail -2
