"""CPU tests of the host-side scheduling logic (no GPU, no compute calls): the band plan, the stripe
rotation over ranks, and -- with two gloo processes -- that the rotation's send/recv pairs match and every
rank trains every stripe exactly once per epoch on the data its neighbour last wrote.
The functions under test are the C-ABI's mfb200_plan_band / mfb200_dist_rotation (include/mfb200.h)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "question-recommendation-system_b200")
sys.path.insert(0, PKG)
import mfb200  # noqa: E402

CONFIGS = [  # BASELINE.json configs (m, n, nnz, k)
    (10_000, 5_000, 1_000_000, 32),
    (138_000, 27_000, 20_000_000, 128),
    (480_000, 17_800, 100_000_000, 128),
    (1_000_000, 625_000, 250_000_000, 128),
]


def coords(p, a, b):
    """Python restatement of band_coord (csrc/kernels.cu): where a rating (a = T row, b = S row) goes."""
    js, bs = divmod(b, p["stripeRows"])
    sb, bl = divmod(bs, p["segS"])
    tb, ai = divmod(a - p["tLo"], p["segT"])
    ga = ai % p["nG"]
    c = sb % p["nC"]
    t = (tb - c * p["S1"]) % p["nTB"]
    return js, sb, bl, c, t, ga


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("world", [1, 2, 8])
def test_plan_band_invariants(cfg, world):
    m, n, nnz, k = cfg
    nS, nT = min(m, n), max(m, n)
    lo_seen = []
    for rank in range(world):
        p = mfb200.plan_band(m, n, nnz, k, world=world, rank=rank)
        assert p["swap_sides"] == (1 if n > m else 0)
        if mfb200.plan_kernel(m, n, nnz, k, world=world, rank=rank) == 6:
            # the item kernel: one pass, every S row of the stripe owned by one (CTA, group), the whole T share of the rank
            # behind one lock word per row; only its two slots per group in shared memory
            k_al = (k + 7) // 8 * 8
            assert 1 <= p["nC"] <= 148 and p["nG"] == p["nWarps"] * 32 // p["L"] and p["nPass"] == 1
            assert p["nStripes"] * p["stripeRows"] >= nS and p["nC"] * p["segS"] >= p["stripeRows"] and p["segS"] < (1 << 13)
            assert p["nTB"] * p["nG"] >= p["segS"] and p["segT"] >= p["tRows"]
            assert p["smem_bytes"] == p["nG"] * 2 * (k_al * 4 + 16)
            lo_seen.append((p["tLo"], p["tRows"]))
            continue
        assert 1 <= p["nC"] <= 148 and p["nG"] == p["nWarps"] * 32 // p["L"] and p["nTB"] == p["nC"] * p["S1"]
        # the S side is covered by stripes x passes x CTAs x band rows, and a band fits in shared memory
        assert p["nStripes"] * p["stripeRows"] >= nS and p["nC"] * p["nPass"] * p["segS"] >= p["stripeRows"]
        # (+ one prefetch slot per group -- row and accumulator pair -- and one dummy S row for the run kernel, k_al <= 128)
        k_al = (k + 7) // 8 * 8
        slots = p["nG"] * (k_al * 4 + 16) + 16 + (k_al * 4 + 12) if (k_al <= 128 and p["L"] == 8) else 0
        # (the cell kernel, picked for small launches, adds one counter per step and four control words)
        extra = p["smem_bytes"] - p["segS"] * (k_al * 4 + 12) - slots
        assert extra in (0, 4 * p["nTB"] + 16) and p["smem_bytes"] <= 232448 - 1024
        assert p["segS"] < (1 << 13)
        # the T side: ranks partition it, bands x groups cover a rank's share
        assert p["tRows"] >= 0 and p["nTB"] * p["segT"] >= p["tRows"] and p["nG"] * p["segT2"] >= p["segT"]
        lo_seen.append((p["tLo"], p["tRows"]))
    assert lo_seen[0][0] == 0 and sum(r for _, r in lo_seen) == nT
    for (lo0, r0), (lo1, _) in zip(lo_seen, lo_seen[1:]):
        assert lo0 + r0 == lo1


def test_band_schedule_is_conflict_free():
    """At every step, the CTAs work on pairwise different T bands, and a T sub-band is handed from CTA c+1
    (step t-S1) to CTA c (step t): what k_sgd_band_epoch's flags rely on."""
    for S1 in (1, 2):
        os.environ["MFB200_RING_S1"] = str(S1)
        try:
            p = mfb200.plan_band(480_000, 17_800, 100_000_000, 128)
        finally:
            os.environ.pop("MFB200_RING_S1")
        nC, nTB = p["nC"], p["nTB"]
        assert p["S1"] == S1
        band = lambda c, t: (c * S1 + t) % nTB  # noqa: E731
        for t in range(0, nTB, 7):
            assert len({band(c, t) for c in range(nC)}) == nC
        for c in range(nC):
            for t in range(S1, nTB, 5):
                assert band(c, t) == band((c + 1) % nC, t - S1)
        # every CTA meets every T band exactly once per pass
        assert sorted(band(3, t) for t in range(nTB)) == list(range(nTB))
    # coordinates: a rating's step is the one at which its CTA meets its T band
    p = mfb200.plan_band(480_000, 17_800, 100_000_000, 128)
    rng = np.random.RandomState(1)
    for a, b in zip(rng.randint(0, 480_000, 200), rng.randint(0, 17_800, 200)):
        js, sb, bl, c, t, ga = coords(p, int(a), int(b))
        assert js == 0 and 0 <= bl < p["segS"] and 0 <= ga < p["nG"] and 0 <= t < p["nTB"]
        assert (c * p["S1"] + t) % p["nTB"] == (int(a) // p["segT"])


@pytest.mark.parametrize("world,spr", [(2, 1), (2, 2), (4, 1), (8, 1), (8, 2)])
def test_rotation_schedule(world, spr):
    ns = spr * world
    for sub in range(2 * ns):
        steps = [mfb200.dist_rotation(world, r, sub, spr) for r in range(world)]
        # all ranks train different stripes; what rank r sends is what rank r-1 receives
        assert len({s["compute"] for s in steps}) == world
        for r, s in enumerate(steps):
            peer = steps[s["send_to"]]
            assert s["send_to"] == (r - 1) % world and peer["recv_from"] == r and peer["recv_stripe"] == s["send_stripe"]
            # the stripe received now is the one trained spr sub-steps later
            later = mfb200.dist_rotation(world, r, sub + spr, spr)
            assert later["compute"] == s["recv_stripe"]
    for r in range(world):  # an epoch: every stripe once; the epoch starts (and therefore ends) on the home stripes
        seen = [mfb200.dist_rotation(world, r, sub, spr)["compute"] for sub in range(ns)]
        assert sorted(seen) == list(range(ns))
        assert seen[:spr] == [spr * r + i for i in range(spr)]


def free_port():
    """A TCP port nobody listens on right now (fixed port numbers collide with leftovers of an earlier run)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


GLOO_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import mfb200
    rank, world, spr, epochs = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(sys.argv[1]), 3
    dist.init_process_group("gloo")
    ns = spr * world
    # every rank holds all stripes; a stripe is (version counter, last writer); only the current holder's copy is valid
    stripes = np.zeros((ns, 2), np.int64)
    trained = np.zeros(ns, np.int64)
    inbox = {}
    for sub in range(epochs * ns):
        s = mfb200.dist_rotation(world, rank, sub, spr)
        if sub >= spr:  # the transfer issued spr sub-steps ago must be here now
            stripes[s["compute"]] = inbox.pop(sub - spr)
        js = s["compute"]
        stripes[js, 0] += 1          # "train": bump the version
        stripes[js, 1] = rank
        trained[js] += 1
        send = torch.from_numpy(stripes[s["send_stripe"]].copy())
        recv = torch.zeros(2, dtype=torch.int64)
        reqs = [dist.isend(send, s["send_to"]), dist.irecv(recv, s["recv_from"])]
        for q in reqs:
            q.wait()
        inbox[sub] = recv.numpy().copy()
        stripes[s["recv_stripe"]] = inbox[sub]
        # what arrived was last written by rank+1 and has been trained once per rank so far in rotation order
        assert recv[1].item() == (rank + 1) %% world, (rank, sub, recv)
    assert (trained == epochs).all(), trained
    # after whole epochs every rank holds its home stripes at version epochs*world
    for i in range(spr):
        assert stripes[spr * rank + i, 0] == epochs * world, (rank, stripes)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(sys.argv[2], "ok%%d" %% rank), "w").write("ok")
""")


@pytest.mark.parametrize("spr", [1, 2])
def test_rotation_with_two_gloo_ranks(tmp_path, spr):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % PKG)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script), str(spr), str(tmp_path)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists(), out.stdout[-3000:]


# ---- sharded load over several ranks: owner of a rating, all-to-all offsets (host logic of Session::load_band) --------
def test_exchange_plan_offsets():
    rng = np.random.RandomState(0)
    for world in (1, 2, 3, 8):
        counts = rng.randint(0, 50, size=(world, world)).astype(np.uint64)
        total = 0
        for me in range(world):
            so, ro, nr = mfb200.dist_exchange_plan(world, me, counts)
            assert so.tolist() == np.concatenate([[0], np.cumsum(counts[me])[:-1]]).tolist()   # my blocks, by destination
            assert ro.tolist() == np.concatenate([[0], np.cumsum(counts[:, me])[:-1]]).tolist()  # arrivals, by source
            assert nr == int(counts[:, me].sum())
            total += nr
        assert total == int(counts.sum())  # every rating arrives exactly once


@pytest.mark.parametrize("cfg", [(480000, 17800, 100_000_000, 128), (138000, 27000, 20_000_000, 128), (300, 70000, 50_000, 32)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_owner_of_a_row_matches_the_planned_bands(cfg, world):
    """The rank a rating is shipped to must be the rank whose T band (plan_band: tLo, tRows) contains its T row."""
    m, n, nnz, k = cfg
    nT = max(m, n)
    plans = [mfb200.plan_band(m, n, nnz, k, world=world, rank=r) for r in range(world)]
    t_seg = plans[0]["tRows"]  # rows per rank = ceil(nT / world): rank 0's band is always full
    rows = np.unique(np.concatenate([np.arange(0, nT, max(1, nT // 997)), [nT - 1], np.arange(world) * t_seg,
                                     np.maximum(np.arange(1, world + 1) * t_seg - 1, 0)]))
    rows = rows[rows < nT]
    for a in rows:
        o = mfb200.dist_owner_of_row(int(a), t_seg, world)
        assert plans[o]["tLo"] <= a < plans[o]["tLo"] + plans[o]["tRows"], (a, o, plans[o])
    assert sum(p["tRows"] for p in plans) == nT


SHARD_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import mfb200
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    m, n, nnz, k = 5000, 300, 40000, 32
    # the caller hands every rank the whole array; a rank reads only its slice (Session::load_band, sharded)
    rng = np.random.RandomState(11)
    R = np.stack([rng.randint(0, m, nnz), rng.randint(0, n, nnz), rng.randint(1, 6, nnz)], 1).astype(np.int64)
    lo, hi = nnz * rank // world, nnz * (rank + 1) // world
    mine = R[lo:hi]
    plan = mfb200.plan_band(m, n, nnz, k, world=world, rank=rank)
    t_seg = mfb200.plan_band(m, n, nnz, k, world=world, rank=0)["tRows"]
    trow = mine[:, 1] if plan["swap_sides"] else mine[:, 0]   # (the permutations are left out: identity maps)
    owner = np.array([mfb200.dist_owner_of_row(int(a), t_seg, world) for a in trow])
    order = np.argsort(owner, kind="stable")                  # the radix sort by destination
    grouped = mine[order]
    cnt = torch.from_numpy(np.bincount(owner, minlength=world).astype(np.int64))
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, cnt)                                # ncclAllGather of the counts
    counts = torch.stack(allc).numpy().astype(np.uint64)
    so, ro, nr = mfb200.dist_exchange_plan(world, rank, counts)
    got = np.zeros((nr, 3), np.int64)
    reqs, landing = [], []
    for q in range(world):                                    # grouped ncclSend / ncclRecv
        sc, rc = int(counts[rank, q]), int(counts[q, rank])
        if q == rank:
            got[ro[q]:ro[q] + rc] = grouped[so[q]:so[q] + sc]
            continue
        if sc:
            reqs.append(dist.isend(torch.from_numpy(grouped[so[q]:so[q] + sc].copy()), q))
        if rc:
            buf = torch.zeros((rc, 3), dtype=torch.int64)
            reqs.append(dist.irecv(buf, q))
            landing.append((buf, int(ro[q]), rc))
    for r_ in reqs:
        r_.wait()
    for buf, off, rc in landing:
        got[off:off + rc] = buf.numpy()
    # what arrived: exactly the ratings of the whole array whose T row lies in this rank's band, each once
    tall = R[:, 1] if plan["swap_sides"] else R[:, 0]
    want = R[(tall >= plan["tLo"]) & (tall < plan["tLo"] + plan["tRows"])]
    key = lambda x: sorted(map(tuple, x.tolist()))
    assert key(got) == key(want), (rank, len(got), len(want))
    tot = torch.tensor([nr])
    dist.all_reduce(tot)
    assert int(tot) == nnz
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(sys.argv[1], "shard_ok%%d" %% rank), "w").write("ok")
""")


def test_sharded_load_exchange_with_two_gloo_ranks(tmp_path):
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER % PKG)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script), str(tmp_path)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert (tmp_path / "shard_ok0").exists() and (tmp_path / "shard_ok1").exists(), out.stdout[-3000:]


def test_planner_picks_the_kernel_and_the_warps_by_ratings_per_cell(monkeypatch):
    """The run kernel with the ring hand-off and 20 warps per CTA where a (group, step) cell holds many ratings (configs
    #2, #3, #4 on one GPU), T-row locks on every SM with 16 warps where it holds few (config #1, #3's share on 4 and 8
    GPUs), the item kernel where a CTA has about one S row per group (#3's share on 2 GPUs: 60 item rows per CTA); the band
    kernel for k > 128; MFB200_KERNEL overrides."""
    for var in ("MFB200_KERNEL", "MFB200_RING_WARPS", "MFB200_RING_CTAS", "MFB200_TLOCK_BELOW", "MFB200_W20_ABOVE"):
        monkeypatch.delenv(var, raising=False)
    c3 = (480_000, 17_800, 100_000_000, 128)
    for cfg, world, kernel, warps in [(c3, 1, 2, 20), (c3, 2, 6, 16), (c3, 4, 5, 16), (c3, 8, 5, 16),
                                      ((138_000, 27_000, 20_000_000, 128), 1, 2, 20), ((10_000, 5_000, 1_000_000, 32), 1, 5, 16),
                                      ((1_000_000, 625_000, 250_000_000, 128), 1, 2, 20)]:
        assert mfb200.plan_kernel(*cfg, world=world) == kernel, (cfg, world)
        p = mfb200.plan_band(*cfg, world=world)
        assert p["nWarps"] == warps and p["nG"] == warps * 4, (cfg, world, p)
        if kernel in (5, 6):
            assert p["nC"] == min(148, p["stripeRows"])  # every SM, whatever the cell size
    assert mfb200.plan_kernel(20_000, 9_000, 3_000_000, 160) == 1  # k > 128: the band kernel, 32 lanes per rating
    monkeypatch.setenv("MFB200_KERNEL", "item")
    assert mfb200.plan_kernel(*c3, world=8) == 6 and mfb200.plan_band(*c3, world=8)["nPass"] == 1
    monkeypatch.setenv("MFB200_KERNEL", "run")
    assert mfb200.plan_kernel(*c3, world=8) == 2
