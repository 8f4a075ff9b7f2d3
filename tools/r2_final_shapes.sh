# usage (GPU box): bash tools/r2_final_shapes.sh -- what the planner picks by default on every shape of interest
for shape in c3 c2 c1 c4 240000,8900,25000000,128 120000,4450,6250000,128 60000,2225,1562500,128; do
  r=$(timeout 300 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas|Error|error" | awk '/^epoch/{printf "%s rmse %s ", $3, $NF} /grid_ctas/{match($0,/.grid_ctas.: [0-9]+/); g=substr($0,RSTART,RLENGTH); match($0,/.cta_warps.: [0-9]+/); w=substr($0,RSTART,RLENGTH); match($0,/.bands.: [0-9]+/); b=substr($0,RSTART,RLENGTH); match($0,/.kernel.: [0-9]+/); k=substr($0,RSTART,RLENGTH); printf "%s %s %s %s", g, w, b, k} /rror/{print}')
  echo "$shape default -> ms(e3,e4)= $r"
done
