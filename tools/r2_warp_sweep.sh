t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s ", $3} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); printf "%s %s", g, b} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
for shape in 60000,2225,1562500,128 120000,4450,6250000,128 240000,8900,25000000,128 c3 c2 c1; do
  t run $shape MFB200_KERNEL=run
  t warp $shape MFB200_KERNEL=warp
  t warp $shape MFB200_KERNEL=warp MFB200_RING_S1=2
  t warp $shape MFB200_KERNEL=warp MFB200_RING_CTAS=148
  t warp $shape MFB200_KERNEL=warp MFB200_RING_CTAS=148 MFB200_RING_S1=2
done
t warp 60000,2225,1562500,128 MFB200_KERNEL=warp MFB200_RING_CTAS=100
t warp 60000,2225,1562500,128 MFB200_KERNEL=warp MFB200_RING_CTAS=120 MFB200_RING_S1=2
t warp 60000,2225,1562500,128 MFB200_KERNEL=warp MFB200_RING_S1=4
MFB200_STATS=1 MFB200_KERNEL=warp python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | grep -E "stats|epoch 3"
MFB200_STATS=1 MFB200_KERNEL=warp python tools/prof_ring.py c3 4 2>&1 | grep -E "stats|epoch 3"
