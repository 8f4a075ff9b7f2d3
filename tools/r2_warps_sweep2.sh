# usage (GPU box): bash tools/r2_warps_sweep2.sh -- the 640-thread build (96 registers) run with 16 / 18 / 20 working warps
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 200 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.cta_warps.: [0-9]+/); w=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); printf "%s %s %s", g, w, b} /rror/{print}')
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
P=question-recommendation-system_b200
for shape in c3 c2 c1 240000,8900,25000000,128 120000,4450,6250000,128 60000,2225,1562500,128 c4; do
  for w in 16 18 20; do t w640 $shape MFB200_LIB=$P/lib_w640/libmf.so MFB200_RING_WARPS=$w; done
done
