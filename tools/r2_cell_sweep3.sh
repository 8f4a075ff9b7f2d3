t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s ", $3} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); printf "%s %s", g, b} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
for shape in 60000,2225,1562500,128 120000,4450,6250000,128; do
  for f in 0 1 2 3; do for s1 in 2 8; do t cell $shape MFB200_KERNEL=cell MFB200_CELL_FLAGS=$f MFB200_CELL_S1=$s1; done; done
  t cell $shape MFB200_KERNEL=cell MFB200_CELL_FLAGS=1 MFB200_CELL_S1=2 MFB200_RING_CTAS=148
  t cell $shape MFB200_KERNEL=cell MFB200_CELL_FLAGS=1 MFB200_CELL_S1=4 MFB200_RING_CTAS=64
done
for f in 0 1; do for s1 in 2 8; do
echo "== stats flags=$f s1=$s1"
MFB200_STATS=1 MFB200_KERNEL=cell MFB200_CELL_FLAGS=$f MFB200_CELL_S1=$s1 python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | grep -E "stats|epoch 3"
done; done
echo "== stats run kernel"
MFB200_STATS=1 MFB200_KERNEL=run python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | grep -E "stats|epoch 3"
