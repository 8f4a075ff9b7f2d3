import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
os.environ["MFB200_KERNEL"] = "warp"
os.environ["MFB200_RING_CTAS"] = "3"
os.environ["MFB200_WAIT_LIMIT_S"] = "2"
import mfb200
m, n, nnz, k, it = 4000, 3000, 300_000, 128, 2
R = mfb200.gen_ratings(m, n, 0, nnz)
s = mfb200.Session(m, n, k, iters=it, mode=mfb200.MODE_RING)
s.load(R)
print(s.report())
try:
    print(s.epochs(it))
except Exception as e:
    print("ERR", e)
