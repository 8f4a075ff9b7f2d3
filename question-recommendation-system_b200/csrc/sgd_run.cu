// csrc/sgd_run.cu -- the throughput SGD kernel, third generation ("run" kernel), sm_100a.
//
// Replaces the per-rating loop SolverBase::run + L2_MFR::prepare_for_sg_update + MFSolver::sg_update
// (mf/mf.cpp:1220-1235, 1720-1728, 1462-1548) and the block scheduler (mf/mf.cpp:113-150, 193-220) for the default
// loss (L2_MFR, no L1 term, no NMF -- what utility_train runs, mf/mf.cpp:4538-4557) and k_al <= 128.  Everything else
// stays on k_sgd_band_epoch (kernels.cu).
//
// Same placement and the same conflict-free schedule as the band kernel: CTA c keeps an S band in
// shared memory, T bands rotate ring-wise over the CTAs, group gamma (8 lanes = one rating) of a CTA owns T sub-band
// gamma of the band it meets at step t, S rows are handed out inside the CTA by locks (or tickets: reproducible).
// What is new -- each item answers a measured limiter of the band kernel (profiles/r1c_band_c3_ncu_full.txt):
//
//   * RUNS.  Inside a (group, step) cell the stream is ordered by T row, so the ratings of one T row are adjacent (a
//     "run", 1.4 ratings on average at the Netflix shape, 26 when the item stripes rotate over 8 GPUs).  The T row and
//     its accumulators stay in REGISTERS for the whole run: one load per run instead of one per rating.
//   * PREFETCH.  While a run is processed the next run's T row travels global -> shared memory by cp.async (LDGSTS,
//     no registers): the first use of a T row no longer waits for L2/HBM (the long-scoreboard stall was 28 % of the
//     band kernel's cycles).  Legal because the row belongs to this group for the whole step -- it is issued only
//     for steps whose hand-off has been verified.
//   * CONVERGED CONTROL.  The four groups of a warp share one instruction stream.  In the band kernel the step
//     hand-off (acquire poll of the neighbour's flag, release fence + flag store) sat in per-group branches, so a
//     warp paid up to four serialised L2 round trips and four fences per iteration -- what bounded small launches
//     (item stripes rotating over several GPUs: ~1 rating per cell, 7 us per step).  Here every group's poll is one
//     predicated load issued one iteration ahead of its use, and one fence per iteration serves every group that
//     acquired or publishes.
//   * SHORT LOCKS.  An S row is locked just before it is read and released as soon as the new row and its
//     accumulators are stored; the T-side half of the update runs outside the lock.
//
// Arithmetic per rating as in the band kernel: z by butterfly shuffle inside the group, e = r - z,
// both gradients from the OLD p and q, G += sum(g^2)/8 for BOTH halves, dims 0-7 and 8..k_al with separate
// accumulators, epoch 0 touches dims 0-7 only, loss in double.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_helpers.cuh"
#include "kernels.h"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int L = 8;   // lanes per group
constexpr int V = 4;   // 16-byte chunks per lane: covers k_al <= 128
constexpr unsigned kNoRow = 0xffffffffu;
constexpr unsigned kBMask = (1u << MFK_W1_BBITS) - 1u;
constexpr unsigned kAMask = (1u << MFK_W0_ABITS) - 1u;

// a 16-byte chunk of a factor row as two packed fp32 pairs: what LDS.128 / LDG.128 deliver and FFMA2 consumes, so no
// instruction is spent on packing
typedef ulonglong2 chunk_t;

// KFULL: k_al == 128, every lane owns four chunks (no per-chunk predicates).  FULL: every dimension is updated
// (false only in epoch 0, which touches dims 0-7: mf/mf.cpp:2834, 2910).
//
// TLK ("T-row locks", args.tlock != NULL; locks only): the ring's step hand-off is replaced by one lock word per T row in
// global memory.  The stream keeps its order -- CTA c starts in T band c*S1 and walks the bands in ring order -- so the CTAs
// are staggered exactly as in the ring, but nobody waits for a whole sub-band: a group takes the lock of the row its next
// run needs (relaxed CAS at L2, one fence per warp iteration for every acquire and release in it, the row then comes
// straight from L2) and gives it back when the run is over.  Still one writer per row and per column at any time
// (mf/mf.cpp:130-142); what goes away is the chain "a T band visits all nC CTAs, every visit ends with a hand-off" that
// bounds a launch with few ratings per cell (item stripes rotating over several GPUs).  A group that waits for a row holds
// no other row, so the waits cannot form a cycle.
//
// The locks go back WITHOUT a fence in the working warps (a fence.acq_rel.gpu per warp iteration cost 0.84 us of the 2.3 us an
// iteration then took): a group whose run is over pushes the row's number into a small queue in shared memory, and one extra
// warp per CTA -- the releaser -- drains the queue: one fence.acq_rel.gpu for everything it found (the working warps' row stores
// happen before their queue entries at CTA scope, the fence is cumulative: the pattern of a grid-wide barrier), then the lock
// words are cleared.  The CAS that takes a lock is relaxed: T rows and their accumulators only ever move through L2
// (ld.global.cg, cp.async.cg, st.global.cg), so there is no stale copy closer to the SM that an acquire would have to drop.
constexpr int kRelQ = 128;  // entries of the release queue (a power of two)

// NW: working warps per CTA the instantiation is compiled for.  16 -> 124 registers per thread; 20 -> 96 registers, no
// spills: five warps per scheduler instead of four hide more of the kernel's latency when a launch has enough ratings per
// cell (C3 18.7 -> 17.9 ms, C2 4.61 -> 4.50); with few ratings per cell the extra groups only add contention for the CTA's S
// rows, and the 96-register code is slower at 16 warps than the 124-register code (profiles/experiments/r2_warps_per_cta.txt).
template <bool DYN, bool STATS, bool KFULL, bool FULL, bool TLK, int NW>
__global__ void __launch_bounds__((NW + (TLK ? 1 : 0)) * 32, 1) k_sgd_run_epoch(const __grid_constant__ mfk_band_args g) {
    // STATS (MFB200_STATS=1): [0] warp iterations, [1] of them with an update, [2] group updates; group-iterations
    // without one because [3] the stream is finished, [4] the T sub-band is not released yet, [5] the S row is busy;
    // [6] runs started from the prefetch slot, [7] runs started with a direct (exposed) load.
    unsigned long long st_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const mfk_band_shape &sh = g.shape;
    const int k_al = g.k_al, nvec = k_al >> 2;
    chunk_t *s_rows = reinterpret_cast<chunk_t *>(smem_raw);                        // [rows_cap + 1][nvec]
    float2 *s_g = reinterpret_cast<float2 *>(s_rows + (sh.rows_cap + 1) * nvec);    // [rows_cap + 1]
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_g + sh.rows_cap + 1);          // [rows_cap + 1]
    // one prefetch slot per group: the row (nvec chunks) + the 16-byte pair of accumulators that contains the row's
    chunk_t *s_slots = reinterpret_cast<chunk_t *>(smem_raw + ((((size_t)(sh.rows_cap + 1) * (nvec * 16 + 12)) + 15) & ~(size_t)15));
    // Row `rows_cap` is a dummy: zeros with accumulators 1.  A group that sits an iteration out computes against it with
    // a zero step, which leaves its T row in registers exactly as it was (p + 0 * g with finite g) -- so the update
    // below needs no per-register predicates.
    const unsigned dummy = (unsigned)sh.rows_cap;

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & (L - 1), gi = lane >> 3;
    const int nG = sh.nG, gamma = warp * 4 + gi;
    const bool leader = l == 0;
    chunk_t *slot = s_slots + gamma * (nvec + 1);

    bool act[V];
#pragma unroll
    for (int j = 0; j < V; j++) act[j] = KFULL || l + L * j < nvec;
    const bool h0 = l < 2;             // chunk 0 of lanes 0,1 = dims 0-7: the first AdaGrad half
    const bool st0 = FULL || h0;       // chunk 0 is stored

    const unsigned nTB = (unsigned)sh.nTB, segT = (unsigned)sh.segT;
    const int S1 = sh.S1;
    const unsigned cS1 = ((unsigned)c * (unsigned)S1) % nTB;
    unsigned *my_flag = g.flags + (size_t)c * nG + gamma;
    const unsigned *nb_flag = g.flags + (size_t)((c + 1) % sh.nC) * nG + gamma;
    const bool ring = !TLK && sh.nC > 1;
    const float eta = g.eta;
    const f32x2 ls2 = pack2(g.lambda_s, g.lambda_s), lt2 = pack2(g.lambda_t, g.lambda_t);
    float *const Tbase = g.T;
    float2 *const TGbase = reinterpret_cast<float2 *>(g.TG);
    double loss = 0.0;
    float lossf = 0.f;  // e*e of the current window, flushed into `loss` (double: mf/mf.cpp:1725-1726) at every refill
    __shared__ int s_dead;
    __shared__ unsigned s_relq[TLK ? kRelQ : 1], s_qtail, s_qhead, s_wdone;
    const bool helper = TLK && warp == sh.nWarps;  // the releaser warp (one more than the plan's working warps)
    if (tid == 0) s_dead = 0;
    for (int i = tid; i < nvec; i += blockDim.x) s_rows[dummy * nvec + i] = make_ulonglong2(0ull, 0ull);
    if (tid == 0) {
        s_g[dummy] = make_float2(1.f, 1.f);
        s_cnt[dummy] = 0u;
    }

    auto t_row = [&](unsigned w0) -> unsigned {
        unsigned tb = cS1 + (w0 >> MFK_W0_ABITS);  // (c*S1 + t) mod nTB without a division: both terms are < nTB
        if (tb >= nTB) tb -= nTB;
        return tb * segT + (w0 & kAMask);
    };

    for (int pass = 0; pass < sh.nPass; ++pass) {
        const int sb = pass * sh.nC + c;
        const int row0 = sb * sh.segS;
        const int nrows = max(0, min(sh.segS, g.nS - row0));
        // ---- stage the S band in ----
        {
            const float4 *src = reinterpret_cast<const float4 *>(g.S) + (size_t)row0 * nvec;
            float4 *dst = reinterpret_cast<float4 *>(s_rows);
            for (int i = tid; i < nrows * nvec; i += blockDim.x) dst[i] = __ldcg(src + i);
            const float2 *srcg = reinterpret_cast<const float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) {
                s_g[i] = __ldcg(srcg + i);
                s_cnt[i] = 0u;
            }
            if (TLK) {
                for (int i = tid; i < kRelQ; i += blockDim.x) s_relq[i] = 0u;
                if (tid == 0) {
                    s_qtail = 0u;
                    s_qhead = 0u;
                    s_wdone = 0u;
                }
            }
        }
        __syncthreads();
        bool dead = false;
        if (helper) {
            // ---- the releaser: entries are taken in order (a prefix of filled slots), one fence per batch ----
            unsigned head = 0u;
            for (;;) {
                const unsigned idx = (head + (unsigned)lane) & (kRelQ - 1);
                const unsigned v = ld_acquire_cta_smem(&s_relq[idx]);
                const unsigned filled = __ballot_sync(kFullMask, v != 0u);
                const int cnt = filled == kFullMask ? 32 : __ffs((int)~filled) - 1;
                if (cnt > 0) {
                    fence_acq_rel_gpu();
                    if (lane < cnt) {
                        st_relaxed_gpu(g.tlock + (v - 1u), 0u);
                        st_volatile_smem(&s_relq[idx], 0u);
                    }
                    __syncwarp();
                    head += (unsigned)cnt;
                    if (lane == 0) st_volatile_smem(&s_qhead, head);
                } else if (*reinterpret_cast<volatile unsigned *>(&s_wdone) == (unsigned)sh.nWarps &&
                           *reinterpret_cast<volatile unsigned *>(&s_qtail) == head) {
                    break;
                }
            }
        } else {

        const unsigned base = g.base + (unsigned)pass * nTB;
        const unsigned done_mark = base + nTB;
        const unsigned pos = g.goff[(size_t)sb * nG + gamma];
        const unsigned end = g.goff[(size_t)sb * nG + gamma + 1];
        unsigned pub = base;  // value of my_flag (all earlier passes / launches are complete)
        // steps <= t_ok have been released to this group by its neighbour (flag >= base + s - S1 + 1); the first S1
        // steps of a launch follow steps of the previous launch, which is complete.  t_cur: the step being worked on.
        int t_ok = !ring ? (int)nTB : (pass == 0 ? S1 - 1 : -1);
        int t_cur = -1;

        // The stream window: two batches of L entries in registers (lane l holds entry l of each batch): the current one
        // and the next one, whose T rows are pulled into L2 half a window before it becomes current.
        // Stream words (kernels.cu, k_band_stream).  Locks: w0 = T row, w1 = step << 13 | S row.  Tickets: w0 = step << 20 |
        // row inside the T band, w1 = ticket << 13 | S row.
        unsigned x0, x1, y0, y1;
        float xr, yr;
        auto ld_batch = [&](unsigned bbase, unsigned &z0, unsigned &z1, float &zr) {
            const unsigned i = bbase + (unsigned)l;
            z0 = 0u; z1 = 0u; zr = 0.f;
            if (i < end) {
                z0 = __ldcs(g.w0 + i);
                z1 = __ldcs(g.w1 + i);
                zr = __ldcs(g.rr + i);
            }
        };
        auto row_of = [&](unsigned w0) -> unsigned { return DYN ? w0 : t_row(w0); };
        auto step_of = [&](unsigned w0, unsigned w1) -> int { return (int)(DYN ? w1 >> MFK_W1_BBITS : w0 >> MFK_W0_ABITS); };
        auto pf_rows = [&](unsigned bbase, unsigned w0) {  // pull the T rows of a batch into L2
            if (bbase + (unsigned)l < end) {
                const unsigned a = row_of(w0);
                const char *rp = reinterpret_cast<const char *>(Tbase + (size_t)a * k_al);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (KFULL || i * 128 < k_al * 4) prefetch_l2(rp + i * 128);
                prefetch_l2(TGbase + a);
            }
        };
        unsigned cbase = pos;
        ld_batch(cbase, x0, x1, xr);
        ld_batch(cbase + L, y0, y1, yr);
        pf_rows(cbase, x0);
        bool pf_next = false;                                            // the next batch's rows have been prefetched
        unsigned nb = cbase < end ? min((unsigned)L, end - cbase) : 0u;  // entries in the current batch
        unsigned hs = 0u;                                                // its first unprocessed entry

        // the run in progress: its T row lives in registers (all lanes), its accumulators in the leader's
        unsigned cur_row = kNoRow, pre_row = kNoRow;  // pre_row: the row whose copy into `slot` has been issued
        chunk_t p[V];
        float2 tg = make_float2(1.f, 1.f);
#pragma unroll
        for (int j = 0; j < V; j++) p[j] = make_ulonglong2(0ull, 0ull);
        // TLK: lk_row = locked, not yet on its way to the slot; cas_row = its CAS was issued in the previous iteration (answer
        // in the leader's cas_old); last_lk = the newest row a lock was obtained for; nx = first window entry (relative to
        // cbase) that no lock covers yet
        unsigned lk_row = kNoRow, cas_row = kNoRow, cas_old = 1u, last_lk = kNoRow;
        int nx = 0;
        unsigned fval = base;  // the neighbour's flag as read one iteration ago
        bool polled = false;
        unsigned idle = 0;
        unsigned long long idle_since = 0;

        for (;;) {
            // (1) current batch used up: the next one becomes current and a new next one is requested
            if (hs == nb && nb != 0u) {
                cbase += L;
                nx = max(nx - L, 0);
                x0 = y0; x1 = y1; xr = yr;
                ld_batch(cbase + L, y0, y1, yr);
                pf_next = false;
                nb = cbase < end ? min((unsigned)L, end - cbase) : 0u;
                hs = 0u;
                loss += (double)lossf;
                lossf = 0.f;
            }
            if (!pf_next && hs >= (unsigned)(L / 2)) {  // (the next batch was requested half a window ago: it is here)
                pf_rows(cbase + L, y0);
                pf_next = true;
            }
            // (2) the stream is walked in order: the head entry and the one after it, broadcast to the group
            const bool valid = hs < nb;
            const unsigned nidx = hs + 1u;
            const bool nin = nidx < (unsigned)L;
            const unsigned hw0 = __shfl_sync(kFullMask, x0, hs, L);
            const unsigned hw1 = __shfl_sync(kFullMask, x1, hs, L);
            const float r = __shfl_sync(kFullMask, xr, hs, L);
            const unsigned nw0 = __shfl_sync(kFullMask, nin ? x0 : y0, nidx & (L - 1), L);
            const unsigned nw1 = __shfl_sync(kFullMask, nin ? x1 : y1, nidx & (L - 1), L);
            const int ht = valid ? step_of(hw0, hw1) : (int)nTB;  // no entry left: "the step after the last"
            const unsigned hrow = row_of(hw0);

            // (3) hand-off between CTAs -- only when a group stands at a step boundary or has a poll to look at
            const bool hand = !TLK && (polled || (valid ? ht != t_cur : pub != done_mark));
            if (!TLK && __any_sync(kFullMask, hand)) {
                // acquiring side: the neighbour's flag as read one iteration ago
                bool acquired = false;
                if (polled) {
                    const int s_rel = (int)(fval - base) + S1 - 1;
                    if (s_rel > t_ok) {
                        t_ok = s_rel;
                        acquired = true;
                    }
                }
                // releasing side: every step before the head's is finished (the stream is walked in order).  A step
                // may be declared complete only after ITS OWN dependency has been verified -- also when the group has
                // no rating in it -- so a group with nothing to do advances step by step, one ahead of its neighbour.
                const unsigned want = base + (unsigned)min(ht, t_ok + 1);
                const bool need_pub = hand && (int)(want - pub) > 0;
                if (__any_sync(kFullMask, acquired || need_pub)) {
                    // one fence per warp iteration: acquire side of every flag consumed above (this lane's T-row loads
                    // below are ordered after it), release side of every flag stored below (this lane's T-row stores
                    // before it)
                    fence_acq_rel_gpu();
                    __syncwarp();
                    if (need_pub) {
                        if (leader) st_relaxed_gpu(my_flag, want);
                        pub = want;
                    }
                }
                if (valid && ht <= t_ok) t_cur = ht;
            }
            bool can = valid && ht == t_cur;
            if (TLK) {
                // (3') T-row locks, pipelined like the row itself: the lock of a run is asked for two runs ahead (relaxed CAS
                // at L2, looked at one iteration later), the locked row then travels to the slot while the run before it is
                // worked on, and the lock goes back when the head has left the row.  Locks are asked for in stream order, so
                // a group that waits for a row only holds rows of EARLIER entries, which it works off without waiting.
                bool acquired = false;
                if (cas_row != kNoRow) {  // (a) the answer to the CAS of the previous iteration (the leader's register)
                    const unsigned old = __shfl_sync(kFullMask, cas_old, 0, L);
                    if (old == 0u) {
                        lk_row = cas_row;
                        last_lk = cas_row;
                        nx++;
                        acquired = true;
                    } else if (STATS && leader) {
                        st_[6]++;
                    }
                    cas_row = kNoRow;
                } else {
                    (void)__shfl_sync(kFullMask, cas_old, 0, L);
                }
                const bool end_run = cur_row != kNoRow && (!valid || hrow != cur_row);  // (b) the run's stores are out
                if (__any_sync(kFullMask, end_run)) {
                    __syncwarp();  // every lane's row stores are ordered before the leader's queue entry
                    if (end_run && leader) {
                        const unsigned t = atomicAdd(&s_qtail, 1u);
                        while ((int)(t - *reinterpret_cast<volatile unsigned *>(&s_qhead)) >= kRelQ) {}
                        st_release_cta_smem(&s_relq[t & (kRelQ - 1)], cur_row + 1u);
                    }
                }
                if (end_run) cur_row = kNoRow;
                (void)acquired;
                const bool swt = valid && cur_row == kNoRow && pre_row == hrow;  // (d) a new run: its row is in the slot
                if (__any_sync(kFullMask, swt)) {
                    cp_async_wait_all();
                    __syncwarp();
                    if (swt) {
#pragma unroll
                        for (int j = 0; j < V; j++)
                            if (act[j]) p[j] = slot[l + L * j];
                        const float4 pair = *reinterpret_cast<const float4 *>(slot + nvec);
                        const bool odd = ((reinterpret_cast<uintptr_t>(TGbase + hrow) >> 3) & 1u) != 0;
                        tg = odd ? make_float2(pair.z, pair.w) : make_float2(pair.x, pair.y);
                        cur_row = hrow;
                        pre_row = kNoRow;
                        if (STATS && leader) st_[7]++;
                    }
                    __syncwarp();
                }
                if (lk_row != kNoRow && pre_row == kNoRow) {  // (e) the locked row starts for the slot
                    const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)lk_row * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) cp_async16(slot + l + L * j, trow + l + L * j);
                    if (leader)
                        cp_async16(slot + nvec, reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(TGbase + lk_row) & ~(uintptr_t)15));
                    cp_async_commit();
                    pre_row = lk_row;
                    lk_row = kNoRow;
                }
                {  // (f) the next entry that no lock covers yet: same row as the last lock, or a CAS for its row
                    const unsigned xr0 = __shfl_sync(kFullMask, nx < L ? x0 : y0, nx & (L - 1), L);
                    if (lk_row == kNoRow && nx < 2 * L && cbase + (unsigned)nx < end) {
                        if (xr0 == last_lk) {
                            nx++;
                        } else {
                            cas_row = xr0;
                            if (leader) cas_old = cas_relaxed_gpu(g.tlock + xr0, 0u, 1u);
                        }
                    }
                }
                can = valid && cur_row == hrow;
            }

            // (4) a new run: its T row comes from the prefetch slot (or, at the start of a step, straight from L2)
            const bool sw = !TLK && can && hrow != cur_row;
            if (!TLK && __any_sync(kFullMask, sw)) {
                cp_async_wait_all();
                __syncwarp();  // the accumulator pair in the slot was copied by the group's leader
                if (sw) {
                    if (pre_row == hrow) {
#pragma unroll
                        for (int j = 0; j < V; j++)
                            if (act[j]) p[j] = slot[l + L * j];
                        const float4 pair = *reinterpret_cast<const float4 *>(slot + nvec);
                        const bool odd = ((reinterpret_cast<uintptr_t>(TGbase + hrow) >> 3) & 1u) != 0;
                        tg = odd ? make_float2(pair.z, pair.w) : make_float2(pair.x, pair.y);
                        pre_row = kNoRow;  // the slot is free again
                        if (STATS && leader) st_[6]++;
                    } else {
                        const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)hrow * k_al);
#pragma unroll
                        for (int j = 0; j < V; j++)
                            if (act[j]) p[j] = __ldcg(trow + l + L * j);
                        tg = __ldcg(TGbase + hrow);
                        if (STATS && leader) st_[7]++;
                    }
                    cur_row = hrow;
                }
                __syncwarp();  // everyone has read the slot before the next copy into it is issued
            }

            // (5) the entry after the head: if it starts another run and its step has been released to this group, its
            // T row starts travelling to the slot now; if its step has not been released, that is what the poll is for
            if (!TLK) {
                const bool nvalid = cbase + nidx < end;
                const unsigned nrow = row_of(nw0);
                const int nt = nvalid ? step_of(nw0, nw1) : ht;
                if (nvalid && nt <= t_ok && nrow != hrow && pre_row == kNoRow) {
                    const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)nrow * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) cp_async16(slot + l + L * j, trow + l + L * j);
                    if (leader)
                        cp_async16(slot + nvec, reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(TGbase + nrow) & ~(uintptr_t)15));
                    cp_async_commit();
                    pre_row = nrow;
                }
                // the poll that the NEXT iteration consumes: needed while a step this group can see is not released yet
                polled = ring && t_ok < min(max(ht, nt), (int)nTB - 1);
                if (polled) fval = ld_relaxed_gpu(nb_flag);
            }

            // (6) everything about the T row that does not need the S row: its squared norms per AdaGrad half (chunk 0 of
            // lanes 0,1 = dims 0-7) and its step sizes eta * rsqrt(G), from the leader's accumulators
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

            // (7) the head's S row: lock (whoever asks first) or ticket (the stream's order: reproducible)
            unsigned got = 0u;
            if (can && leader) {
                unsigned *cnt = &s_cnt[hw1 & kBMask];
                if (DYN)
                    got = cas_acquire_cta_smem(cnt, 0u, 1u) == 0u;
                else
                    got = (ld_acquire_cta_smem(cnt) & MFK_TICKET_MASK) == (hw1 >> MFK_W1_BBITS);
            }
            got = __shfl_sync(kFullMask, got, 0, L);
            const bool ready = got != 0u;
            const unsigned bl = ready ? (hw1 & kBMask) : dummy;
            if (STATS) {
                if (lane == 0) st_[0]++;
                if (leader) {
                    if (ready) st_[2]++;
                    else if (!valid) st_[3]++;
                    else if (!can) st_[4]++;
                    else st_[5]++;
                }
            }

            if (!__any_sync(kFullMask, ready)) {
                if (__all_sync(kFullMask, nb == 0u && (TLK ? (cur_row == kNoRow && pre_row == kNoRow && lk_row == kNoRow && cas_row == kNoRow)
                                                               : pub == done_mark)))
                    break;
                // A wait that never ends (a lost hand-off would be a bug; a dead neighbour GPU is not): give up after
                // a generous wall-clock limit so that the kernel terminates, and tell the other warps and CTAs.
                if (++idle >= 4096u) {
                    idle = 0;
                    const unsigned long long now = global_timer_ns();
                    if (idle_since == 0) idle_since = now;
                    if (now - idle_since > g.wait_limit_ns || *reinterpret_cast<volatile int *>(g.error_flag) != 0) {
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

            // ---- the update (sg_update, mf/mf.cpp:1462-1548, 1228-1234), executed by all groups of the warp.  A group
            // that is not ready works on the dummy row with a zero step: nothing of it is stored, its registers keep
            // their values.
            // With g_s = lambda_s q - e p and g_t = lambda_t p - e q (both from the OLD rows, mf/mf.cpp:1476-1491):
            //   q' = q - eta_s g_s = (1 - eta_s lambda_s) q + (eta_s e) p          (two packed operations per pair)
            //   sum g_s^2 = lambda_s^2 <q,q> - 2 lambda_s e <p,q> + e^2 <p,p>       (no second reduction)
            // so one round of shuffles -- <p,q> and <q,q> per AdaGrad half -- serves the error, both new rows and all
            // four accumulators, and the S row is held for: load, 16 FFMA2, 3 shuffle stages, 16 packed operations,
            // store. ----
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
            // S side: new row, new accumulators, then the row is free
            {
                const float es0 = eta * rsqrtf(sg.x), es1 = FULL ? eta * rsqrtf(sg.y) : 0.f;
                const float esa = h0 ? es0 : es1;  // chunk 0 of lanes 0,1 belongs to the first half
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
                    // G += sum(g^2) / 8 for BOTH halves (the shipped SSE path's rk)
                    float2 sgn = sg;
                    sgn.x += fmaf(l2, qq0, fmaf(m2, pq0, e2 * pp0)) * 0.125f;
                    if (FULL) sgn.y += fmaf(l2, qq_all - qq0, fmaf(m2, pq_all - pq0, e2 * (pp_all - pp0))) * 0.125f;
                    s_g[bl] = sgn;
                }
                __syncwarp();  // the group's shared-memory stores are ordered before the release of the row
                if (ready && leader)
                    st_release_cta_smem(&s_cnt[bl], DYN ? 0u : ((hw1 >> MFK_W1_BBITS) + 1u) & MFK_TICKET_MASK);
            }
            // T side: the row belongs to this group for the whole step; the new row stays in registers for the rest of
            // the run and is written through (the next group to own it may run on another SM)
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
                    hs++;
                }
            }
        }
        loss += (double)lossf;
        lossf = 0.f;
        if (TLK) {  // this warp's last releases are in the queue
            __syncwarp();
            if (lane == 0) atomicAdd(&s_wdone, 1u);
        }
        }  // (working warps)

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
        if (any_dead) break;  // CTA-uniform: no warp goes on to a pass its siblings have left
    }

    if (!leader) loss = 0.0;  // every lane of a group accumulated the group's e*e
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

// shared memory the run kernel needs beside the S band: one prefetch slot (row + accumulator pair) per group
unsigned mfk_sgd_run_slot_bytes(int k_al, int groups) {
    // + the dummy S row (row, accumulators, lock word) that groups sitting out an iteration compute against
    return (unsigned)groups * (unsigned)(k_al * 4 + 16) + 16u + (unsigned)(k_al * 4 + 12);
}

int mfk_sgd_run_max_warps(void) { return 20; }

int mfk_sgd_run_supported(int k_al, int L_, int fun, float lambda1_s, float lambda1_t, int do_nmf) {
    return k_al <= 128 && L_ == 8 && fun == MFK_FUN_L2_MFR && lambda1_s == 0.f && lambda1_t == 0.f && !do_nmf;
}

int mfk_sgd_run_epoch(const mfk_band_args *args, void *stream) {
    const bool st = args->stats != nullptr, dy = args->dynamic != 0, kf = args->k_al == 128;
    if (!mfk_sgd_run_supported(args->k_al, args->shape.L, args->fun, args->lambda1_s, args->lambda1_t, args->do_nmf))
        return (int)cudaErrorInvalidValue;
    const bool fu = args->full != 0, tl = args->tlock != nullptr, w20 = args->shape.nWarps > 16;
    if (args->shape.nWarps > 20) return (int)cudaErrorInvalidValue;
    if (tl && !dy) return (int)cudaErrorInvalidValue;  // T-row locks are a timing-dependent order: not for the ticket mode
#define MFB_RUN3(D, S, K, F, T) (w20 ? (const void *)k_sgd_run_epoch<D, S, K, F, T, 20> : (const void *)k_sgd_run_epoch<D, S, K, F, T, 16>)
#define MFB_RUN2(D, S, K, T) (fu ? MFB_RUN3(D, S, K, true, T) : MFB_RUN3(D, S, K, false, T))
#define MFB_RUN(D, S, T) (kf ? MFB_RUN2(D, S, true, T) : MFB_RUN2(D, S, false, T))
    const void *fn = tl   ? (st ? MFB_RUN(true, true, true) : MFB_RUN(true, false, true))
                     : dy ? (st ? MFB_RUN(true, true, false) : MFB_RUN(true, false, false))
                          : (st ? MFB_RUN(false, true, false) : MFB_RUN(false, false, false));
#undef MFB_RUN3
#undef MFB_RUN2
#undef MFB_RUN
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)args->shape.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block((args->shape.nWarps + (tl ? 1 : 0)) * 32);  // (+ the releaser warp)
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, args->shape.smem_bytes, (cudaStream_t)stream);
}

}  // extern "C"
