t() { local label=$1 shape=$2; shift 2
  r=$(env "$@" timeout 300 python tools/prof_ring.py $shape 5 2>&1 | grep -E "^epoch [34]|grid_ctas" | awk '/^epoch/{printf "%s ", $3} /grid_ctas/{match($0,/.bands.: [0-9]+/); printf "%s", substr($0,RSTART,RLENGTH)}')
  echo "$shape $label $* -> $r"; }
for s1 in 2 3 4; do t s1 c3 MFB200_RING_S1=$s1; done
for s1 in 1 2; do t s1 c2 MFB200_RING_S1=$s1; t s1 240000,8900,25000000,128 MFB200_RING_S1=$s1; done
for s1 in 1 2; do t s1 c4 MFB200_RING_S1=$s1; done
