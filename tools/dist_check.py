#!/usr/bin/env python
"""Multi-GPU check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [workload] [epochs] [nnz]
Trains with the item half-stripes rotating over N ranks and prints, on rank 0, one JSON line with the held-out
RMSE, the per-epoch device time (max over ranks) and whether every rank ended with the same model.  Exits non-zero
when the held-out RMSE of the model on the device and of the model handed to the host differ, or when ranks disagree."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import bench  # noqa: E402
import mfb200  # noqa: E402

rank, world, local = (int(os.environ.get(x, d)) for x, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m, n, nnz, k, desc = bench.WORKLOADS[wl]
if len(sys.argv) > 3:
    nnz = int(sys.argv[3])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.from_numpy(mfb200.dist_unique_id()))
dist.broadcast(idt, 0)
R = mfb200.gen_ratings(m, n, 0, nnz)
# held-out ratings as in tests/golden/named_configs.json (the compiled reference's values at equal epochs live there)
T = mfb200.gen_ratings(m, n, nnz, min(nnz // 10, 10_000_000))
golden = {}
if len(sys.argv) <= 3:
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "named_configs.json"))).get(wl, {}).get("runs", {})
s = mfb200.Session(m, n, k, iters=epochs, rank=rank, world=world, nccl_id=idt.cpu().numpy(), lam_p=bench.LAMBDA,
                   lam_q=bench.LAMBDA, eta=bench.ETA, mode=mfb200.MODE_RING, device=local)
s.load(R)
times, trs = [], []
for e in range(epochs):
    torch.cuda.synchronize()
    dist.barrier()
    ms, tr = s.epochs(1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
    trs.append(float(tr[0]))
rm = s.rmse(T)
P, Q, b = s.finish()
rep = s.report()
s.close()
h = hashlib.sha256(P.tobytes() + Q.tobytes()).digest()[:8]
ht = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
hs = [torch.zeros_like(ht) for _ in range(world)]
dist.all_gather(hs, ht)
same = all(bool((x == hs[0]).all()) for x in hs)
rc = 0
if rank == 0:
    host_rm = mfb200.rmse(T, P, Q, b)
    # the model that reaches the caller (finish -> host arrays -> re-upload) must be the model on the device:
    # same kernel, same ratings, so the two sums are equal to the last bit
    if abs(rm - host_rm) >= 1e-9 or not same:
        print("dist_check: FAILED heldout_rmse %.12f (device model) vs %.12f (host model), all_ranks_same_model=%s"
              % (rm, host_rm, same), file=sys.stderr, flush=True)
        rc = 1
    ref = golden.get(str(epochs), {}).get("heldout_rmse")
    print(json.dumps({"workload": desc, "nnz": nnz, "world": world, "epochs": epochs, "ms_per_epoch": times,
                      "rmse_parity": {"ours": rm, "reference": ref, "epochs": epochs,
                                      "rel": None if ref is None else rm / ref - 1,
                                      "ok": None if ref is None else bool(abs(rm / ref - 1) < 0.005)},
                      "updates_per_s_last": nnz / times[-1] * 1e3, "tr_rmse": trs, "heldout_rmse": rm,
                      "heldout_rmse_host_model": host_rm, "all_ranks_same_model": same, "ok": rc == 0,
                      "schedule": {x: rep[x] for x in ("grid_ctas", "cta_warps", "bands", "subbands", "launches")}}),
          flush=True)
dist.destroy_process_group()
sys.exit(rc)
