# usage (GPU box): bash tools/r2_topk_stride.sh -- top-k on random and on TRAINED factors: centring level (0 none, 1 items, 2 both
# sides + item bias + items sorted by bias; default: chosen from the factors), pass-A stride
run() { echo "== $*"; env "$@" MFB200_TOPK_STATS=1 python tools/bench_topk.py $ARGS 2>&1 | grep -E "topk stats: [0-9.]+ cand|users_per_s|rror" | cut -c1-330 | tail -2; }
ARGS="75776 500000 128 100 2"
run AUTO=1
ARGS="75776 625000 128 100 2 trained=10"
run AUTO=1
run MFB200_TOPK_STRIDE=3
ARGS="75776 625000 128 100 2 trained=20"
run AUTO=1
