// csrc/sgd_warp.cu -- the throughput SGD kernel with WARP-owned T sub-bands ("warp" kernel), sm_100a.
//
// Same job, placement, arithmetic and hand-off protocol as k_sgd_run_epoch (sgd_run.cu): the per-rating loop
// SolverBase::run + L2_MFR::prepare_for_sg_update + MFSolver::sg_update (mf/mf.cpp:1220-1235, 1720-1728, 1462-1548)
// under a schedule in which no two concurrent updates share a row or a column (the guarantee of the reference's block
// scheduler, mf/mf.cpp:130-142); default loss (L2_MFR, no L1 term, no NMF), k_al <= 128, S rows handed out by locks.
//
// What changes is the unit that owns a T sub-band for a step: the WARP instead of the group (16 sub-bands per T band
// instead of 64).  Measured on the run kernel when the item stripes rotate over 8 GPUs (a launch of 1.56M ratings:
// ~4 ratings per (group, step) cell): only 1.9 of a warp's 4 groups have work in an iteration, because every group waits
// for its own small, Poisson-sized cell (profiles/r2_run_vs_band_shapes.txt, profiles/experiments/r2_cell_kernel.txt).
// Here the four groups of a warp are served from ONE stream:
//
//   * the warp holds a window of 64 consecutive entries of its stream (two batches of 32, one entry per lane) and a
//     cursor `head`: everything before it has been handed to a group;
//   * the ratings of one T row inside a cell are adjacent (a run) and a run is the unit that is handed out -- its
//     first entry is where the T row changes -- so two groups never hold the same T row; a group that finishes a run
//     gets the next run after the cursor, whatever cell it is in, as long as that step has been released;
//   * every group holds its current run and one run ahead, whose T row travels global -> shared memory by cp.async
//     while the current run is worked on (as in the run kernel);
//   * one flag per (CTA, warp): 4x fewer polls, fences and flag stores, cells 4x larger, and a step without ratings
//     costs the warp nothing but the flag.
//
// The groups share one instruction stream, so all the bookkeeping above is warp-uniform register arithmetic
// (ballots, find-first-set, shuffles): no shared-memory atomics, no counters.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dev_helpers.cuh"
#include "kernels.h"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int L = 8;   // lanes per group
constexpr int V = 4;   // 16-byte chunks per lane: covers k_al <= 128
constexpr unsigned kNoRow = 0xffffffffu;
constexpr unsigned kBMask = (1u << MFK_W1_BBITS) - 1u;
constexpr int kNone = -1;  // "no entry"
constexpr int W = 32;      // entries per batch of the window

typedef ulonglong2 chunk_t;

template <bool STATS, bool KFULL, bool FULL>
__global__ void __launch_bounds__(512, 1) k_sgd_warp_epoch(const __grid_constant__ mfk_band_args g) {
    // STATS (MFB200_STATS=1): [0] warp iterations, [1] of them with an update, [2] group updates; group-iterations
    // without one because [3] the group has no run (stream finished or nothing released), [4] unused, [5] the S row is
    // busy; [6] runs started from the prefetch slot, [7] runs started with a direct (exposed) load.
    unsigned long long st_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const mfk_band_shape &sh = g.shape;
    const int k_al = g.k_al, nvec = k_al >> 2;
    chunk_t *s_rows = reinterpret_cast<chunk_t *>(smem_raw);                        // [rows_cap + 1][nvec]
    float2 *s_g = reinterpret_cast<float2 *>(s_rows + (sh.rows_cap + 1) * nvec);    // [rows_cap + 1]
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_g + sh.rows_cap + 1);          // [rows_cap + 1]
    chunk_t *s_slots = reinterpret_cast<chunk_t *>(smem_raw + ((((size_t)(sh.rows_cap + 1) * (nvec * 16 + 12)) + 15) & ~(size_t)15));
    const unsigned dummy = (unsigned)sh.rows_cap;  // zero row with accumulators 1: what a group sitting out computes against

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & (L - 1), gi = lane >> 3;
    const int nU = sh.nG;  // ownership units per CTA = warps
    const bool leader = l == 0;
    chunk_t *slot = s_slots + (warp * 4 + gi) * (nvec + 1);

    bool act[V];
#pragma unroll
    for (int j = 0; j < V; j++) act[j] = KFULL || l + L * j < nvec;
    const bool h0 = l < 2;
    const bool st0 = FULL || h0;

    const int nTB = sh.nTB, S1 = sh.S1;
    unsigned *my_flag = g.flags + (size_t)c * nU + warp;
    const unsigned *nb_flag = g.flags + (size_t)((c + 1) % sh.nC) * nU + warp;
    const bool ring = sh.nC > 1;
    const float eta = g.eta;
    float *const Tbase = g.T;
    float2 *const TGbase = reinterpret_cast<float2 *>(g.TG);
    double loss = 0.0;
    float lossf = 0.f;
    __shared__ int s_dead;
    if (tid == 0) s_dead = 0;
    for (int i = tid; i < nvec; i += blockDim.x) s_rows[dummy * nvec + i] = make_ulonglong2(0ull, 0ull);
    if (tid == 0) {
        s_g[dummy] = make_float2(1.f, 1.f);
        s_cnt[dummy] = 0u;
    }

    for (int pass = 0; pass < sh.nPass; ++pass) {
        const int sb = pass * sh.nC + c;
        const int row0 = sb * sh.segS;
        const int nrows = max(0, min(sh.segS, g.nS - row0));
        {   // ---- stage the S band in ----
            const float4 *src = reinterpret_cast<const float4 *>(g.S) + (size_t)row0 * nvec;
            float4 *dst = reinterpret_cast<float4 *>(s_rows);
            for (int i = tid; i < nrows * nvec; i += blockDim.x) dst[i] = __ldcg(src + i);
            const float2 *srcg = reinterpret_cast<const float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) {
                s_g[i] = __ldcg(srcg + i);
                s_cnt[i] = 0u;
            }
        }
        __syncthreads();

        const unsigned base = g.base + (unsigned)pass * (unsigned)nTB;
        const unsigned done_mark = base + (unsigned)nTB;
        const unsigned pos0 = g.goff[(size_t)sb * nU + warp];
        const unsigned end = g.goff[(size_t)sb * nU + warp + 1];
        unsigned pub = base;
        int t_ok = !ring ? nTB : (pass == 0 ? S1 - 1 : -1);

        // the window: entries cb + [0, 64) of the warp's stream; lane j holds entries cb + j (x) and cb + 32 + j (y).
        // Stream words with locks (kernels.cu, k_band_stream): w0 = T row, w1 = step << 13 | S row.
        unsigned cb = pos0;
        unsigned x0 = kNoRow, x1 = 0u, y0 = kNoRow, y1 = 0u;
        float xr = 0.f, yr = 0.f;
        auto ld_batch = [&](unsigned bbase, unsigned &z0, unsigned &z1, float &zr) {
            const unsigned i = bbase + (unsigned)lane;
            z0 = kNoRow; z1 = 0u; zr = 0.f;
            if (i < end) {
                z0 = __ldcs(g.w0 + i);
                z1 = __ldcs(g.w1 + i);
                zr = __ldcs(g.rr + i);
            }
        };
        ld_batch(cb, x0, x1, xr);
        ld_batch(cb + W, y0, y1, yr);
        unsigned prev_row = kNoRow;          // T row of the entry before the window
        unsigned long long heads = 0ull;     // bit i: entry cb + i starts a run
        bool heads_stale = true;
        int head = 0;                        // first entry of the window not yet handed to a group
        // fetch the words of window entry `at` (0..63; any value per lane)
        auto fetch = [&](int at, unsigned &z0, unsigned &z1, float &zr) {
            const int src = at & (W - 1);
            const unsigned a0 = __shfl_sync(kFullMask, x0, src), b0 = __shfl_sync(kFullMask, y0, src);
            const unsigned a1 = __shfl_sync(kFullMask, x1, src), b1 = __shfl_sync(kFullMask, y1, src);
            const float ar = __shfl_sync(kFullMask, xr, src), br = __shfl_sync(kFullMask, yr, src);
            const bool lo = at < W;
            z0 = lo ? a0 : b0;
            z1 = lo ? a1 : b1;
            zr = lo ? ar : br;
        };

        // per group: the run at work (cur) and the run handed out ahead of it (nxt)
        int cur_at = kNone, nxt_at = kNone;  // window positions of the entry at work / of the next run's first entry
        bool cur_known = false;              // the words of the entry at cur_at have been fetched
        unsigned cw0 = kNoRow, cw1 = 0u, nw0 = kNoRow, nw1 = 0u;
        float cr = 0.f, nr = 0.f;
        unsigned cur_row = kNoRow;
        chunk_t p[V];
        float2 tg = make_float2(1.f, 1.f);
#pragma unroll
        for (int j = 0; j < V; j++) p[j] = make_ulonglong2(0ull, 0ull);
        unsigned fval = base;
        bool polled = false;
        unsigned idle = 0;
        unsigned long long idle_since = 0;
        bool dead = false;

        for (;;) {
            // (1) the window moves on by one batch when nothing in its first half is needed any more
            {
                const bool lowfree = (cur_at == kNone || cur_at >= W) && (nxt_at == kNone || nxt_at >= W);
                if (head >= W && __all_sync(kFullMask, lowfree)) {
                    prev_row = __shfl_sync(kFullMask, x0, W - 1);
                    x0 = y0; x1 = y1; xr = yr;
                    cb += W;
                    ld_batch(cb + W, y0, y1, yr);
                    head -= W;
                    if (cur_at != kNone) cur_at -= W;
                    if (nxt_at != kNone) nxt_at -= W;
                    heads_stale = true;
                    loss += (double)lossf;
                    lossf = 0.f;
                }
            }
            const int nvalid = (int)min((unsigned)(2 * W), end - cb);  // entries of the window that exist
            if (heads_stale) {
                unsigned px = __shfl_up_sync(kFullMask, x0, 1), py = __shfl_up_sync(kFullMask, y0, 1);
                const unsigned xl = __shfl_sync(kFullMask, x0, W - 1);
                if (lane == 0) {
                    px = prev_row;
                    py = xl;
                }
                const unsigned hx = __ballot_sync(kFullMask, lane < nvalid && x0 != px);
                const unsigned hy = __ballot_sync(kFullMask, lane + W < nvalid && y0 != py);
                heads = (unsigned long long)hx | ((unsigned long long)hy << 32);
                heads_stale = false;
                // the cursor never rests on the continuation of a run (it belongs to the group that was handed the run's
                // first entry, possibly a window ago): on to the next run head, or to the end of what is known
                if (head < nvalid && ((heads >> head) & 1ull) == 0ull) {
                    const unsigned long long later = heads & ~((1ull << head) - 1ull);
                    head = later ? __ffsll((long long)later) - 1 : nvalid;
                }
            }

            // (2) hand-off, acquiring side: the neighbour's flag as polled one iteration ago
            bool acquired = false;
            if (polled) {
                const int s_rel = (int)(fval - base) + S1 - 1;
                if (s_rel > t_ok) {
                    t_ok = s_rel;
                    acquired = true;
                }
            }
            // entries whose step has been released to this warp: a prefix of the window (the stream is ordered by step)
            const unsigned rx = __ballot_sync(kFullMask, lane < nvalid && (int)(x1 >> MFK_W1_BBITS) <= t_ok);
            const unsigned ry = __ballot_sync(kFullMask, lane + W < nvalid && (int)(y1 >> MFK_W1_BBITS) <= t_ok);
            const unsigned long long released = (unsigned long long)rx | ((unsigned long long)ry << 32);

            // (3) every group: go on inside the run, or begin the run that was handed out ahead
            if (cur_at != kNone && !cur_known) {
                // the entry after the one just finished: the same run unless a new one starts there or the stream ends
                if (cur_at >= nvalid) {
                    if (cb + (unsigned)cur_at >= end) cur_at = kNone;  // (else: beyond the window, wait for it to move)
                } else if ((heads >> cur_at) & 1ull) {
                    cur_at = kNone;
                }
            }
            bool begin = false;
            if (cur_at == kNone && nxt_at != kNone) {
                cur_at = nxt_at;
                cw0 = nw0; cw1 = nw1; cr = nr;
                nxt_at = kNone;
                cur_known = true;
                begin = true;
            }
            {
                unsigned f0, f1;
                float fr;
                fetch(cur_at == kNone ? 0 : cur_at, f0, f1, fr);
                if (cur_at != kNone && !cur_known && cur_at < nvalid) {
                    cw0 = f0; cw1 = f1; cr = fr;
                    cur_known = true;
                }
            }

            // (3b) a run begins: its T row is in the slot (before the next copy into the slot is issued in (4))
            if (__any_sync(kFullMask, begin)) {
                cp_async_wait_all();
                __syncwarp();
                if (begin) {
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) p[j] = slot[l + L * j];
                    const float4 pair = *reinterpret_cast<const float4 *>(slot + nvec);
                    const bool odd = ((reinterpret_cast<uintptr_t>(TGbase + cw0) >> 3) & 1u) != 0;
                    tg = odd ? make_float2(pair.z, pair.w) : make_float2(pair.x, pair.y);
                    cur_row = cw0;
                    if (STATS && leader) st_[6]++;
                }
                __syncwarp();  // everyone has read the slot before the next copy into it is issued
            }

            // (4) hand out runs: every group without a run ahead gets the next run after the cursor whose step has been
            // released; its T row starts travelling to the group's slot
            {
                const unsigned want = __ballot_sync(kFullMask, leader && nxt_at == kNone);  // bits 0, 8, 16, 24
                unsigned long long avail = head < 2 * W ? heads & released & ~((1ull << head) - 1ull) : 0ull;
                int mine = kNone;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (((want >> (8 * q)) & 1u) && avail) {
                        const int at = __ffsll((long long)avail) - 1;
                        avail &= avail - 1ull;
                        if (q == gi) mine = at;
                        // the cursor moves past this run: to the next run head, released or not, or to the end of what is known
                        const unsigned long long later = heads & ~((2ull << at) - 1ull);
                        head = later ? __ffsll((long long)later) - 1 : nvalid;
                    }
                }
                unsigned f0, f1;
                float fr;
                fetch(mine == kNone ? 0 : mine, f0, f1, fr);
                // a group without a run at work begins the run at once (its T row comes straight from L2: one exposed
                // load, but no iteration is lost); the others take it as the run ahead, T row by cp.async into the slot
                const bool direct = mine != kNone && cur_at == kNone;
                if (mine != kNone && !direct) {
                    nxt_at = mine;
                    nw0 = f0; nw1 = f1; nr = fr;
                    const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)nw0 * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) cp_async16(slot + l + L * j, trow + l + L * j);
                    if (leader)
                        cp_async16(slot + nvec, reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(TGbase + nw0) & ~(uintptr_t)15));
                    cp_async_commit();
                }
                if (direct) {
                    cur_at = mine;
                    cw0 = f0; cw1 = f1; cr = fr;
                    cur_known = true;
                    const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)cw0 * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) p[j] = __ldcg(trow + l + L * j);
                    tg = __ldcg(TGbase + cw0);
                    cur_row = cw0;
                    if (STATS && leader) st_[7]++;
                }
            }

            // (5) hand-off, releasing side.  Every entry of a step earlier than `low` is done: entries are handed out in
            // order, so what is not done is at the cursor or behind it, or is a group's run at work / ahead.  A step may be
            // declared complete only after ITS OWN dependency has been verified.
            {
                int low = nTB;
                if (head < nvalid) {
                    unsigned h0w, h1w;
                    float hr;
                    fetch(head, h0w, h1w, hr);
                    low = (int)(h1w >> MFK_W1_BBITS);
                } else if (cb + (unsigned)nvalid < end) {  // more entries beyond the window: not before the last step seen
                    unsigned h0w, h1w;
                    float hr;
                    fetch(nvalid - 1, h0w, h1w, hr);
                    low = (int)(h1w >> MFK_W1_BBITS);
                }
                // (cw1: the entry at work -- or, while the next entry of the run is still beyond the window, the entry
                // just finished, which belongs to the same run and so to the same step)
                int mine = nTB;
                if (cur_at != kNone) mine = (int)(cw1 >> MFK_W1_BBITS);
                if (nxt_at != kNone) mine = min(mine, (int)(nw1 >> MFK_W1_BBITS));
                low = min(low, __reduce_min_sync(kFullMask, mine));
                const unsigned want = base + (unsigned)min(low, t_ok + 1);
                const bool need_pub = (int)(want - pub) > 0;
                if (acquired || need_pub) {  // warp-uniform
                    // one fence per warp iteration: acquire side of the flag consumed above (this lane's T-row loads are
                    // ordered after it), release side of the flag stored below (this lane's T-row stores before it)
                    fence_acq_rel_gpu();
                    __syncwarp();
                    if (need_pub) {
                        if (lane == 0) st_relaxed_gpu(my_flag, want);
                        pub = want;
                    }
                }
                // the poll the NEXT iteration consumes: while an entry of the window waits for its step, or the stream is
                // through and steps without ratings are left to declare
                const bool blocked = nvalid > 0 && ((released >> (nvalid - 1)) & 1ull) == 0ull;
                const bool tail = cb + (unsigned)nvalid >= end && pub != done_mark;
                polled = ring && t_ok < nTB - 1 && (blocked || tail);
                if (polled) {
                    unsigned v = 0u;
                    if (lane == 0) v = ld_relaxed_gpu(nb_flag);
                    fval = __shfl_sync(kFullMask, v, 0);
                }
            }

            // (7) what does not need the S row: squared norms of the T row per AdaGrad half, its step sizes
            float pp_all, pp0;
            {
                f32x2 na = mul2(p[0].x, p[0].x);
                na = fma2(p[0].y, p[0].y, na);
                f32x2 nbv = pack2(0.f, 0.f);
#pragma unroll
                for (int j = 1; j < V; j++) {
                    nbv = fma2(p[j].x, p[j].x, nbv);
                    nbv = fma2(p[j].y, p[j].y, nbv);
                }
                const float c0n = sum2(na);
                pp0 = h0 ? c0n : 0.f;
                pp_all = c0n + sum2(nbv);
                pp0 += __shfl_xor_sync(kFullMask, pp0, 1);
#pragma unroll
                for (int o = L / 2; o > 0; o >>= 1) pp_all += __shfl_xor_sync(kFullMask, pp_all, o);
            }
            const float et0 = __shfl_sync(kFullMask, eta * rsqrtf(tg.x), 0, L);
            const float et1 = FULL ? __shfl_sync(kFullMask, eta * rsqrtf(tg.y), 0, L) : 0.f;

            // (8) the S row of the entry at work: whoever asks first
            const bool can = cur_at != kNone && cur_known;
            unsigned got = 0u;
            if (can && leader) got = cas_acquire_cta_smem(&s_cnt[cw1 & kBMask], 0u, 1u) == 0u;
            got = __shfl_sync(kFullMask, got, 0, L);
            const bool ready = got != 0u;
            const unsigned bl = ready ? (cw1 & kBMask) : dummy;
            const float r = cr;
            if (STATS) {
                if (lane == 0) st_[0]++;
                if (leader) {
                    if (ready) st_[2]++;
                    else if (!can) st_[3]++;
                    else st_[5]++;
                }
            }

            if (!__any_sync(kFullMask, ready)) {
                const bool none_left = cur_at == kNone && nxt_at == kNone;
                if (__all_sync(kFullMask, none_left) && head >= nvalid && cb + (unsigned)nvalid >= end && pub == done_mark) break;
                // nothing to work on and a poll outstanding: the warp stays on its neighbour's flag until it moves (a poll
                // per loop iteration would add the loop's own latency to every hand-off)
                if (polled && !__any_sync(kFullMask, can)) {
                    unsigned v = fval;
                    for (int sp = 0; sp < 24; sp++) {
                        if (lane == 0) v = ld_relaxed_gpu(nb_flag);
                        v = __shfl_sync(kFullMask, v, 0);
                        if (v != fval) break;
                    }
                    fval = v;
                }
                if (++idle >= 4096u) {
                    idle = 0;
                    const unsigned long long now = global_timer_ns();
                    if (idle_since == 0) idle_since = now;
                    if (now - idle_since > g.wait_limit_ns || *reinterpret_cast<volatile int *>(g.error_flag) != 0) {
#ifdef MFB_WARP_DEBUG
                        if (leader)
                            printf("dead c=%d w=%d g=%d pass=%d head=%d nvalid=%d cb=%u pos0=%u end=%u t_ok=%d pub=%u base=%u cur=%d known=%d nxt=%d cw1step=%d heads=%llx rel=%llx fval=%u polled=%d\n",
                                   c, warp, gi, pass, head, nvalid, cb, pos0, end, t_ok, pub, base, cur_at, (int)cur_known, nxt_at,
                                   (int)(cw1 >> MFK_W1_BBITS), heads, released, fval, (int)polled);
#endif
                        if (lane == 0) atomicCAS(g.error_flag, 0, 2);
                        dead = true;
                        break;
                    }
                }
                continue;
            }
            idle = 0;
            idle_since = 0;
            if (STATS && lane == 0) st_[1]++;

            // ---- the update (sg_update, mf/mf.cpp:1462-1548, 1228-1234); see sgd_run.cu for the algebra ----
            chunk_t *srow = s_rows + bl * nvec;
            chunk_t q[V];
#pragma unroll
            for (int j = 0; j < V; j++) q[j] = act[j] ? srow[l + L * j] : make_ulonglong2(0ull, 0ull);
            const float2 sg = s_g[bl];
            float pq_all, pq0, qq_all, qq0;
            {
                f32x2 da = mul2(p[0].x, q[0].x), qa = mul2(q[0].x, q[0].x);
                da = fma2(p[0].y, q[0].y, da);
                qa = fma2(q[0].y, q[0].y, qa);
                f32x2 db = pack2(0.f, 0.f), qb = db;
#pragma unroll
                for (int j = 1; j < V; j++) {
                    db = fma2(p[j].x, q[j].x, db);
                    qb = fma2(q[j].x, q[j].x, qb);
                    db = fma2(p[j].y, q[j].y, db);
                    qb = fma2(q[j].y, q[j].y, qb);
                }
                const float d0 = sum2(da), q0 = sum2(qa);
                pq0 = h0 ? d0 : 0.f;
                qq0 = h0 ? q0 : 0.f;
                pq_all = d0 + sum2(db);
                qq_all = q0 + sum2(qb);
                pq0 += __shfl_xor_sync(kFullMask, pq0, 1);
                qq0 += __shfl_xor_sync(kFullMask, qq0, 1);
#pragma unroll
                for (int o = L / 2; o > 0; o >>= 1) {
                    pq_all += __shfl_xor_sync(kFullMask, pq_all, o);
                    qq_all += __shfl_xor_sync(kFullMask, qq_all, o);
                }
            }
            const float e = r - pq_all;  // mf/mf.cpp:1724 (z = <p,q>, calc_z 1264-1273)
            const float gate = ready ? 1.f : 0.f;
            lossf = fmaf(gate * e, e, lossf);
            {
                const float es0 = eta * rsqrtf(sg.x), es1 = FULL ? eta * rsqrtf(sg.y) : 0.f;
                const float esa = h0 ? es0 : es1;
                const float k1 = fmaf(-es1, g.lambda_s, 1.f), k2 = es1 * e, ka1 = fmaf(-esa, g.lambda_s, 1.f), ka2 = esa * e;
                const f32x2 k1v = pack2(k1, k1), k2v = pack2(k2, k2), ka1v = pack2(ka1, ka1), ka2v = pack2(ka2, ka2);
#pragma unroll
                for (int j = 0; j < V; j++) {
                    chunk_t qn;
                    qn.x = fma2(j == 0 ? ka2v : k2v, p[j].x, mul2(j == 0 ? ka1v : k1v, q[j].x));
                    qn.y = fma2(j == 0 ? ka2v : k2v, p[j].y, mul2(j == 0 ? ka1v : k1v, q[j].y));
                    if (ready && act[j] && (j == 0 ? st0 : FULL)) srow[l + L * j] = qn;
                }
                if (ready && leader) {
                    const float ls = g.lambda_s, m2 = -2.f * ls * e, e2 = e * e, l2 = ls * ls;
                    float2 sgn = sg;
                    sgn.x += fmaf(l2, qq0, fmaf(m2, pq0, e2 * pp0)) * 0.125f;
                    if (FULL) sgn.y += fmaf(l2, qq_all - qq0, fmaf(m2, pq_all - pq0, e2 * (pp_all - pp0))) * 0.125f;
                    s_g[bl] = sgn;
                }
                __syncwarp();
                if (ready && leader) st_release_cta_smem(&s_cnt[bl], 0u);
            }
            {
                const float eg0 = gate * et0, eg1 = gate * et1;
                const float eta_a = h0 ? eg0 : eg1;
                const float k1 = fmaf(-eg1, g.lambda_t, 1.f), k2 = eg1 * e, ka1 = fmaf(-eta_a, g.lambda_t, 1.f), ka2 = eta_a * e;
                const f32x2 k1v = pack2(k1, k1), k2v = pack2(k2, k2), ka1v = pack2(ka1, ka1), ka2v = pack2(ka2, ka2);
                chunk_t *trow = reinterpret_cast<chunk_t *>(Tbase + (size_t)cur_row * k_al);
#pragma unroll
                for (int j = 0; j < V; j++) {
                    p[j].x = fma2(j == 0 ? ka2v : k2v, q[j].x, mul2(j == 0 ? ka1v : k1v, p[j].x));
                    p[j].y = fma2(j == 0 ? ka2v : k2v, q[j].y, mul2(j == 0 ? ka1v : k1v, p[j].y));
                    if (ready && act[j] && (j == 0 ? st0 : FULL)) __stcg(trow + l + L * j, p[j]);
                }
                if (ready) {
                    if (leader) {
                        const float lt = g.lambda_t, m2 = -2.f * lt * e, e2 = e * e, l2 = lt * lt;
                        tg.x += fmaf(l2, pp0, fmaf(m2, pq0, e2 * qq0)) * 0.125f;
                        if (FULL) tg.y += fmaf(l2, pp_all - pp0, fmaf(m2, pq_all - pq0, e2 * (qq_all - qq0))) * 0.125f;
                        __stcg(TGbase + cur_row, tg);
                    }
                    // this entry is done; the group looks at the next entry of the window at the top of the loop
                    cur_at += 1;
                    cur_known = false;
                }
            }
        }
        loss += (double)lossf;
        lossf = 0.f;

        // ---- stage the S band out ----
        if (dead) s_dead = 1;
        __syncthreads();
        {
            float4 *dst = reinterpret_cast<float4 *>(g.S) + (size_t)row0 * nvec;
            const float4 *src = reinterpret_cast<const float4 *>(s_rows);
            for (int i = tid; i < nrows * nvec; i += blockDim.x) __stcg(dst + i, src[i]);
            float2 *dstg = reinterpret_cast<float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) __stcg(dstg + i, s_g[i]);
        }
        const int any_dead = s_dead;
        __syncthreads();
        if (any_dead) break;
    }

    if (!leader) loss = 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(kFullMask, loss, o);
    if (lane == 0 && loss != 0.0) atomicAdd(g.loss, loss);
    if (STATS && g.stats) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long v = st_[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
            if (lane == 0 && v) atomicAdd(g.stats + i, v);
        }
    }
}

}  // namespace

extern "C" {

int mfk_sgd_warp_epoch(const mfk_band_args *args, void *stream) {
    const bool st = args->stats != nullptr, kf = args->k_al == 128, fu = args->full != 0;
    if (!mfk_sgd_run_supported(args->k_al, args->shape.L, args->fun, args->lambda1_s, args->lambda1_t, args->do_nmf) ||
        !args->dynamic || args->shape.by_row != 3 || args->shape.nG != args->shape.nWarps)
        return (int)cudaErrorInvalidValue;
#define MFB_WARP2(S, K) (fu ? (const void *)k_sgd_warp_epoch<S, K, true> : (const void *)k_sgd_warp_epoch<S, K, false>)
    const void *fn = st ? (kf ? MFB_WARP2(true, true) : MFB_WARP2(true, false)) : (kf ? MFB_WARP2(false, true) : MFB_WARP2(false, false));
#undef MFB_WARP2
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)args->shape.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block(args->shape.nWarps * 32);
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, args->shape.smem_bytes, (cudaStream_t)stream);
}

}  // extern "C"
