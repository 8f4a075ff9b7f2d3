#!/usr/bin/env python
"""Repeats the multi-rank end-to-end call of bench.py (session create, load from the host array, K epochs, finish) in every
rank of a torchrun job and prints the phases of every call: where does the preprocessing time of several ranks go?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_repeat_dist.py [epochs] [calls]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import bench  # noqa: E402
import mfb200  # noqa: E402

rank, world, local = (int(os.environ.get(x, d)) for x, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m, n, nnz, k, _ = bench.WORKLOADS["c3"]
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R = mfb200.gen_ratings(m, n, 0, nnz)


def new_id():
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.from_numpy(mfb200.dist_unique_id()))
    dist.broadcast(idt, 0)
    return idt.cpu().numpy()


for c in range(calls):
    nid = new_id()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("---- call %d" % c, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    s = mfb200.Session(m, n, k, iters=K, rank=rank, world=world, nccl_id=nid, lam_p=bench.LAMBDA, lam_q=bench.LAMBDA,
                       eta=bench.ETA, mode=mfb200.MODE_RING, device=local)
    t1 = time.perf_counter()
    s.load(R)
    t2 = time.perf_counter()
    s.epochs(K)
    t3 = time.perf_counter()
    s.finish(download=(rank == 0))
    t4 = time.perf_counter()
    s.close()
    t5 = time.perf_counter()
    print("rank %d call %d: create %.1f load %.1f epochs %.1f finish %.1f close %.1f ms" % (
        rank, c, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3), flush=True)
dist.destroy_process_group()
