# one rank's share of C3 on 8 / 4 / 2 GPUs (and C3, C2 themselves) timed alone on one GPU: run kernel vs cell kernel
# usage: bash tools/r2_cell_sweep.sh > gpurun_out/r2_cell_sweep.txt
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s ", $3} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); printf "%s %s", g, b} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
for shape in 60000,2225,1562500,128 120000,4450,6250000,128 240000,8900,25000000,128; do
  t run $shape MFB200_KERNEL=run
  t cell $shape MFB200_KERNEL=cell
  for c in 64 80 100 124 148; do t cell $shape MFB200_KERNEL=cell MFB200_RING_CTAS=$c; done
  for s1 in 2 4 16; do t cell $shape MFB200_KERNEL=cell MFB200_CELL_S1=$s1; done
  for ch in 1 2 4 8; do t cell $shape MFB200_KERNEL=cell MFB200_CELL_CHUNK=$ch; done
done
for shape in c3 c2 c1; do
  t run $shape MFB200_KERNEL=run
  t cell $shape MFB200_KERNEL=cell
  t cell $shape MFB200_KERNEL=cell MFB200_CELL_S1=4
done
MFB200_STATS=1 MFB200_KERNEL=cell python tools/prof_ring.py 60000,2225,1562500,128 4 2>&1 | tail -3
MFB200_STATS=1 MFB200_KERNEL=cell python tools/prof_ring.py c3 4 2>&1 | tail -3
