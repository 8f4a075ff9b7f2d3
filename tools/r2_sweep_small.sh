for shape in 60000,2225,1562500,128 120000,4450,6250000,128; do
for s1 in 1 2 4; do for c in 78 110 148; do
  r=$(MFB200_RING_S1=$s1 MFB200_RING_CTAS=$c timeout 100 python tools/prof_ring.py $shape 5 2>&1 | grep -E "epoch 4" | awk '{print $3}')
  echo "$shape S1=$s1 ctas=$c ms=$r"
done; done; done
