# usage (GPU box): bash tools/r2_warps_sweep.sh -- run kernel compiled for 16 / 18 / 20 / 21 working warps per CTA
# (make VARIANT=w640 EXTRA=-DMFB_RUN_THREADS=640 ...): more warps per scheduler against fewer registers per thread
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.cta_warps.: [0-9]+/); w=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); printf "%s %s %s", g, w, b} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
P=question-recommendation-system_b200
for shape in c3 c2 c1 240000,8900,25000000,128 120000,4450,6250000,128 60000,2225,1562500,128; do
  t w16 $shape
  for v in 576 640 672; do t w$v $shape MFB200_LIB=$P/lib_w$v/libmf.so; done
done
for s1 in 1 2; do t w640_s1 c3 MFB200_LIB=$P/lib_w640/libmf.so MFB200_RING_S1=$s1; done
MFB200_STATS=1 MFB200_LIB=$P/lib_w640/libmf.so python tools/prof_ring.py c3 4 2>&1 | grep -E "stats|epoch 3" | tail -2
