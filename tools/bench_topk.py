#!/usr/bin/env python
"""Top-k scoring bench (BASELINE.json config #5 shape): python tools/bench_topk.py [users] [items] [k] [topk] [reps] [trained=E]
trained=E: the factors are not random but TRAINED by the engine -- E epochs on the Yahoo-R1-shape ratings (1M x 625k, 250M;
`items` is then 625000) -- so that the candidate rate is measured on the score distribution of a real model.
Prints one JSON line: users/s of the scoring (device time, factors resident), the same through the C-ABI call with
host buffers, the fraction of the measured bf16 tensor peak (2*n*k flop per user), and a parity
check of a user sample against the oracle (bit-exact indices)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200  # noqa: E402

rank, world, local = (int(os.environ.get(x, d)) for x, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
if world > 1:  # one process per GPU: users are sharded, every rank holds all item factors; no data-path collective
    os.environ["MFB200_DEVICE"] = str(local)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nusers = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 128
topk = int(sys.argv[4]) if len(sys.argv) > 4 else 100
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
trained = int(sys.argv[6].split("=")[1]) if len(sys.argv) > 6 and sys.argv[6].startswith("trained=") else 0
m = nusers
b_model = 3.5
if trained:
    import bench
    m, n, nnz_t, k, _ = bench.WORKLOADS["c4"]
    Rt = mfb200.gen_ratings(m, n, 0, nnz_t)
    P, Q, b_model, _ = mfb200.train(Rt, m, n, k, trained, lam_p=bench.LAMBDA, lam_q=bench.LAMBDA, eta=bench.ETA,
                                    mode=mfb200.MODE_RING, device=local)
    del Rt
else:
    rng = np.random.RandomState(5)
    P = (rng.rand(m, k).astype(np.float32) * 0.35 + rng.standard_normal((m, k)).astype(np.float32) * 0.1)
    Q = (rng.rand(n, k).astype(np.float32) * 0.35 + rng.standard_normal((n, k)).astype(np.float32) * 0.1)
users = np.arange(nusers, dtype=np.int32)
mine = users[rank * nusers // world:(rank + 1) * nusers // world]  # this rank's users
best_dev, best_wall = 1e30, 1e30
for r in range(reps):
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    t0 = time.perf_counter()
    idx_l, sc_l = mfb200.topk(P, Q, b_model, mine, topk)
    wall, dev = time.perf_counter() - t0, mfb200.topk_last_ms() * 1e-3
    if world > 1:  # the job's time is the slowest rank's
        t = torch.tensor([wall, dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev = float(t[0]), float(t[1])
    best_wall, best_dev = min(best_wall, wall), min(best_dev, dev)
idx = np.full((nusers, topk), -2, np.int32)
sc = np.zeros((nusers, topk), np.float32)
idx[mine], sc[mine] = idx_l, sc_l
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] * 1e12
    src = "measured bf16_tflops_sustained"
except Exception:
    peak, src = 1.4e15, "fallback"
out = {"metric": "topk_users_per_sec", "n_gpus": world, "factors": ("trained: %d epochs at the 1M x 625k shape" % trained) if trained else "random", "users": nusers, "items": n, "k": k, "topk": topk,
       "device_seconds": best_dev, "users_per_s": nusers / best_dev, "e2e_seconds_host_buffers": best_wall,
       "e2e_users_per_s": nusers / best_wall, "algorithmic_flop_per_user": 2.0 * n * k,
       "tensor_roofline": {"achieved_tflops_per_gpu": nusers / best_dev * 2.0 * n * k / 1e12 / world,
                           "peak_tflops": peak / 1e12, "frac": nusers / best_dev * 2.0 * n * k / peak / world,
                           "peak_source": src, "note": "per GPU; the pipeline does ~1.5 GEMM passes per user"}}
if os.path.exists(os.path.join(ROOT, "oracle", "libmf_oracle.so")):
    import orc
    samp = mine[np.linspace(0, len(mine) - 1, 16).astype(np.int64)]
    t0 = time.perf_counter()
    io, so = orc.oracle_topk(P, Q, b_model, samp, topk)
    cpu_s = time.perf_counter() - t0
    out["parity_sample"] = {"users": len(samp), "indices_bit_exact": bool(np.array_equal(idx[samp], io)),
                            "scores_bit_exact": bool(np.array_equal(sc[samp].view(np.uint32), so.view(np.uint32)))}
    out["cpu_baseline"] = {"value": len(samp) / cpu_s, "unit": "users/s", "cores": 1, "kind": "port",
                           "sample": "%d users of the same shape, mf_predict loop + partial sort" % len(samp)}
if world > 1:
    ok = torch.tensor([1 if out.get("parity_sample", {}).get("indices_bit_exact", True) else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["all_ranks_parity"] = bool(ok.item())
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps(out))
