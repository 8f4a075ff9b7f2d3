# usage (GPU box): bash tools/r2_rq_sweep.sh -- ring form with 20 working warps (fences in the working warps) against 19 working
# warps + a releaser warp that publishes the step hand-off (no fence in the working warps)
t() {  # label, shape, env...
  local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 300 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.cta_warps.: [0-9]+/); w=substr($0,RSTART,RLENGTH); match($0,/.kernel.: [0-9]+/); k=substr($0,RSTART,RLENGTH); printf "%s %s %s", g, w, k} /rror/{print}' | cut -c1-200)
  echo "$shape $label $* -> ms(e3,e4)= $r"
}
for shape in c3 c2 240000,8900,25000000,128 c4; do
  t w20 $shape MFB200_KERNEL=run
  t w19rq $shape MFB200_KERNEL=run MFB200_RING_WARPS=19
done
t w19rq_s1 c3 MFB200_KERNEL=run MFB200_RING_WARPS=19 MFB200_RING_S1=1
t w19rq_s3 c3 MFB200_KERNEL=run MFB200_RING_WARPS=19 MFB200_RING_S1=3
