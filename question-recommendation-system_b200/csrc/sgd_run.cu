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
constexpr int V = 4;   // float4 chunks per lane: covers k_al <= 128
constexpr unsigned kNoRow = 0xffffffffu;
constexpr unsigned kBMask = (1u << MFK_W1_BBITS) - 1u;

template <bool DYN, bool STATS>
__global__ void __launch_bounds__(512, 1) k_sgd_run_epoch(const __grid_constant__ mfk_band_args g) {
    // STATS (MFB200_STATS=1): [0] warp iterations, [1] of them with an update, [2] group updates; group-iterations
    // without one because [3] the stream is finished, [4] the T sub-band is not released yet, [5] the S row is busy;
    // [6] runs started from the prefetch slot, [7] runs started with a direct (exposed) load.
    unsigned long long st_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const mfk_band_shape &sh = g.shape;
    const int k_al = g.k_al, nvec = k_al >> 2;
    float4 *s_rows = reinterpret_cast<float4 *>(smem_raw);                          // [rows_cap][nvec]
    float2 *s_g = reinterpret_cast<float2 *>(s_rows + (size_t)sh.rows_cap * nvec);  // [rows_cap]
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_g + sh.rows_cap);              // [rows_cap]
    // one prefetch slot per group: the row (nvec float4) + the 16-byte pair of accumulators that contains the row's
    float4 *s_slots = reinterpret_cast<float4 *>(smem_raw + ((((size_t)sh.rows_cap * (nvec * 16 + 12)) + 15) & ~(size_t)15));

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & (L - 1), gi = lane >> 3, gshift = gi * L;
    const int nG = sh.nG, gamma = warp * 4 + gi;
    const bool leader = l == 0;
    const bool full = g.full != 0;  // slow_only == false
    float4 *slot = s_slots + (size_t)gamma * (nvec + 1);

    bool act[V];
#pragma unroll
    for (int j = 0; j < V; j++) act[j] = l + L * j < nvec;
    const bool h0 = l < 2;  // chunk 0 of lanes 0,1 = dims 0-7: the first AdaGrad half

    auto gballot = [&](bool pr) -> unsigned { return (__ballot_sync(kFullMask, pr) >> gshift) & 0xffu; };

    const unsigned nTB = (unsigned)sh.nTB;
    const unsigned cS1 = ((unsigned)c * (unsigned)sh.S1) % nTB;
    unsigned *my_flag = g.flags + (size_t)c * nG + gamma;
    const unsigned *nb_flag = g.flags + (size_t)((c + 1) % sh.nC) * nG + gamma;
    const bool ring = sh.nC > 1;
    double loss = 0.0;
    __shared__ int s_dead;
    if (tid == 0) s_dead = 0;

    for (int pass = 0; pass < sh.nPass; ++pass) {
        const int sb = pass * sh.nC + c;
        const int row0 = sb * sh.segS;
        const int nrows = max(0, min(sh.segS, g.nS - row0));
        // ---- stage the S band in ----
        {
            const float4 *src = reinterpret_cast<const float4 *>(g.S) + (size_t)row0 * nvec;
            for (int i = tid; i < nrows * nvec; i += blockDim.x) s_rows[i] = __ldcg(src + i);
            const float2 *srcg = reinterpret_cast<const float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) {
                s_g[i] = __ldcg(srcg + i);
                s_cnt[i] = 0u;
            }
        }
        __syncthreads();

        const unsigned base = g.base + (unsigned)pass * nTB;
        const unsigned done_mark = base + nTB;
        const unsigned pos = g.goff[(size_t)sb * nG + gamma];
        const unsigned end = g.goff[(size_t)sb * nG + gamma + 1];
        unsigned pub = base;  // value of my_flag (all earlier passes / launches are complete)
        // steps <= t_ok have been released to this group by its neighbour (flag >= base + s - S1 + 1); the first S1
        // steps of a launch follow steps of the previous launch, which is complete
        int t_ok = !ring ? (int)nTB : (pass == 0 ? sh.S1 - 1 : -1);

        // three batches of L stream entries in registers: lane l holds entry (batch base + l)
        unsigned c0, c1, n0, n1, m0, m1;
        float cr, nr, mr;
        auto ld_batch = [&](unsigned bbase, unsigned &x0, unsigned &x1, float &xr) {
            const unsigned i = bbase + (unsigned)l;
            x0 = 0u; x1 = 0u; xr = 0.f;
            if (i < end) {
                x0 = __ldcs(g.w0 + i);
                x1 = __ldcs(g.w1 + i);
                xr = __ldcs(g.rr + i);
            }
        };
        auto t_row = [&](unsigned w0) -> unsigned {
            const unsigned t = w0 >> MFK_W0_ABITS, ai = w0 & ((1u << MFK_W0_ABITS) - 1u);
            unsigned tb = cS1 + t;  // (c*S1 + t) mod nTB without a division: both terms are < nTB
            if (tb >= nTB) tb -= nTB;
            return tb * (unsigned)sh.segT + ai;
        };
        auto pf_rows = [&](unsigned bbase, unsigned x0) {  // pull the T rows of a batch into L2
            if (bbase + (unsigned)l < end) {
                const unsigned a = t_row(x0);
                prefetch_l2_bulk(g.T + (size_t)a * k_al, (unsigned)k_al * 4u);
                prefetch_l2(g.TG + 2 * (size_t)a);
            }
        };
        unsigned cbase = pos;
        ld_batch(cbase, c0, c1, cr);
        ld_batch(cbase + L, n0, n1, nr);
        ld_batch(cbase + 2 * L, m0, m1, mr);
        pf_rows(cbase, c0);
        pf_rows(cbase + L, n0);
        int nb = cbase < end ? (int)min((unsigned)L, end - cbase) : 0;  // entries in the window
        unsigned done = 0u;                                             // bit i: entry i of the window is processed
        unsigned my_row = nb > l ? t_row(c0) : kNoRow;                  // T row of this lane's window entry

        // the run in progress: its T row and accumulators live in registers
        unsigned cur_row = kNoRow, pre_row = kNoRow;  // pre_row: the row whose copy into `slot` has been issued
        float4 p[V];
        float2 tg = make_float2(1.f, 1.f);
#pragma unroll
        for (int j = 0; j < V; j++) p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned fval = base;      // the neighbour's flag as read one iteration ago
        bool polled = false;
        unsigned idle = 0;
        unsigned long long idle_since = 0;
        bool dead = false;

        for (;;) {
            // (1) window used up: promote the next batch
            if (nb > 0 && done == ((1u << nb) - 1u)) {
                c0 = n0; c1 = n1; cr = nr;
                n0 = m0; n1 = m1; nr = mr;
                cbase += L;
                ld_batch(cbase + 2 * L, m0, m1, mr);
                pf_rows(cbase + L, n0);
                nb = cbase < end ? (int)min((unsigned)L, end - cbase) : 0;
                done = 0u;
                my_row = nb > l ? t_row(c0) : kNoRow;
            }
            // (2) the oldest pending entry (the stream is walked in order, run by run)
            const bool pend = l < nb && !((done >> l) & 1u);
            const unsigned pb = gballot(pend);
            const int hsel = pb ? __ffs(pb) - 1 : 0;
            const unsigned hw0 = __shfl_sync(kFullMask, c0, hsel, L);
            const unsigned hrow = __shfl_sync(kFullMask, my_row, hsel, L);
            const int ht = pb ? (int)(hw0 >> MFK_W0_ABITS) : (int)nTB;  // nothing pending here means the stream is finished

            // (3) hand-off, acquiring side: the flag value read one iteration ago
            bool acquired = false;
            if (polled) {
                const int s_rel = (int)(fval - base) + sh.S1 - 1;
                if (s_rel > t_ok) {
                    t_ok = s_rel;
                    acquired = true;
                }
            }
            // hand-off, releasing side: every step before the oldest pending entry's is finished; a step may be
            // declared complete only after ITS OWN dependency has been verified -- also when the group has no rating
            // in it -- so a group with nothing to do advances step by step, one ahead of its neighbour
            const unsigned want = base + (unsigned)min(ht, t_ok + 1);
            const bool need_pub = (int)(want - pub) > 0;
            if (__any_sync(kFullMask, acquired || need_pub)) {
                // one fence per warp iteration: acquire side of every flag read above (this lane's T-row loads below
                // are ordered after it) and release side of every flag store below (this lane's T-row stores before it)
                fence_acq_rel_gpu();
                __syncwarp();
                if (need_pub) {
                    if (leader) st_relaxed_gpu(my_flag, want);
                    pub = want;
                }
            }

            // (4) the run of the oldest pending entry; switch rows when it is not the row in registers
            const bool can = pb != 0u && ht <= t_ok;
            const bool sw = can && hrow != cur_row;
            if (__any_sync(kFullMask, sw)) {
                cp_async_wait_all();
                __syncwarp();  // the accumulator pair in the slot was copied by the group's leader
                if (sw) {
                    if (pre_row == hrow) {
#pragma unroll
                        for (int j = 0; j < V; j++)
                            if (act[j]) p[j] = slot[l + L * j];
                        const float4 pair = slot[nvec];
                        const bool odd = ((reinterpret_cast<uintptr_t>(g.TG + 2 * (size_t)hrow) >> 3) & 1u) != 0;
                        tg = odd ? make_float2(pair.z, pair.w) : make_float2(pair.x, pair.y);
                        if (STATS && leader) st_[6]++;
                    } else {
                        const float4 *trow = reinterpret_cast<const float4 *>(g.T + (size_t)hrow * k_al);
#pragma unroll
                        for (int j = 0; j < V; j++)
                            if (act[j]) p[j] = __ldcg(trow + l + L * j);
                        tg = __ldcg(reinterpret_cast<const float2 *>(g.TG) + hrow);
                        if (STATS && leader) st_[7]++;
                    }
                    cur_row = hrow;
                    pre_row = kNoRow;  // the slot is free again
                }
                __syncwarp();  // everyone has read the slot before the next copy into it is issued
            }

            // (5) candidates: the pending entries of the run; take one whose S row is available
            // (tickets: the entries of a run are taken oldest first, so the order of updates of the T row is the stream's
            // and a run is reproducible bit for bit; locks: any entry of the run whose S row is free)
            const bool mine = can && pend && my_row == cur_row && (DYN || l == hsel);
            bool elig = false;
            if (mine) {
                const unsigned cnt = ld_acquire_cta_smem(&s_cnt[c1 & kBMask]);
                elig = DYN ? cnt == 0u : (cnt & MFK_TICKET_MASK) == (c1 >> MFK_W1_BBITS);
            }
            unsigned eb = gballot(elig);
            const int sel = eb ? __ffs(eb) - 1 : 0;
            const unsigned x1 = __shfl_sync(kFullMask, c1, sel, L);
            const float r = __shfl_sync(kFullMask, cr, sel, L);
            const unsigned bl = x1 & kBMask;
            if (DYN) {  // the row looked free: try to take its lock (another group may have been faster)
                unsigned got = 0u;
                if (eb && leader) got = cas_acquire_cta_smem(&s_cnt[bl], 0u, 1u) == 0u;
                got = __shfl_sync(kFullMask, got, 0, L);
                if (!got) eb = 0u;
            }
            const bool ready = eb != 0u;

            // (6) the next run's T row: first pending entry of another row, in this window or at the head of the next
            // batch.  Its step must have been released to this group; if not, that is what the next poll is for.
            {
                const bool other = pend && my_row != cur_row;
                const unsigned ob = gballot(other);
                const int nsel = ob ? __ffs(ob) - 1 : 0;
                unsigned nw0 = __shfl_sync(kFullMask, c0, nsel, L);
                unsigned nrow = __shfl_sync(kFullMask, my_row, nsel, L);
                bool have_next = ob != 0u;
                const unsigned bw0 = __shfl_sync(kFullMask, n0, 0, L);
                if (!have_next && cbase + L < end) {
                    nw0 = bw0;
                    nrow = t_row(bw0);
                    have_next = nrow != cur_row;
                }
                const int nt = have_next ? (int)(nw0 >> MFK_W0_ABITS) : ht;
                if (have_next && nt <= t_ok && pre_row == kNoRow && nrow != cur_row) {
                    const float4 *trow = reinterpret_cast<const float4 *>(g.T + (size_t)nrow * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) cp_async16(slot + l + L * j, trow + l + L * j);
                    if (leader) {
                        const uintptr_t a16 = reinterpret_cast<uintptr_t>(g.TG + 2 * (size_t)nrow) & ~(uintptr_t)15;
                        cp_async16(slot + nvec, reinterpret_cast<const void *>(a16));
                    }
                    cp_async_commit();
                    pre_row = nrow;
                }
                // the poll that the NEXT iteration consumes: needed while a step this group can see is not released yet
                const int look = max(ht, nt);
                polled = ring && t_ok < min(look, (int)nTB - 1);
                if (polled) fval = ld_relaxed_gpu(nb_flag);
            }
            if (STATS) {
                if (lane == 0) st_[0]++;
                if (leader) {
                    if (ready) st_[2]++;
                    else if (!pb) st_[3]++;
                    else if (!can) st_[4]++;
                    else st_[5]++;
                }
            }

            if (!__any_sync(kFullMask, ready)) {
                if (__all_sync(kFullMask, nb == 0 && pub == done_mark)) break;
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
                    __nanosleep(64);
                }
                continue;
            }
            idle = 0;
            idle_since = 0;
            if (STATS && lane == 0) st_[1]++;

            // ---- the update, executed by all groups of the warp; only ready groups commit ----
            float4 *srow = s_rows + (size_t)bl * nvec;
            float4 q[V];
#pragma unroll
            for (int j = 0; j < V; j++)
                q[j] = (ready && act[j]) ? srow[l + L * j] : make_float4(0.f, 0.f, 0.f, 0.f);
            float2 sg = ready ? s_g[bl] : make_float2(1.f, 1.f);

            // z = <p,q>  (calc_z, mf/mf.cpp:1264-1273); packed fp32: one FFMA2 covers two dimensions
            f32x2 pp[V][2], qq[V][2];
            float part;
            {
                f32x2 part2 = pack2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < V; j++) {
                    pp[j][0] = pack2(p[j].x, p[j].y);
                    pp[j][1] = pack2(p[j].z, p[j].w);
                    qq[j][0] = pack2(q[j].x, q[j].y);
                    qq[j][1] = pack2(q[j].z, q[j].w);
                    part2 = fma2(pp[j][0], qq[j][0], part2);
                    part2 = fma2(pp[j][1], qq[j][1], part2);
                }
                part = sum2(part2);
            }
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) part += __shfl_xor_sync(kFullMask, part, o);
            const float e = r - part;                          // mf/mf.cpp:1724
            if (ready && leader) loss += (double)(e * e);      // mf/mf.cpp:1725-1726
            const f32x2 ne2 = pack2(-e, -e);

            // sg_update (mf/mf.cpp:1462-1548, 1228-1234), S side first: new row, new accumulators, then the row is free
            {
                const float eta_s0 = g.eta * rsqrtf(sg.x), eta_s1 = g.eta * rsqrtf(sg.y);
                const f32x2 ls2 = pack2(g.lambda_s, g.lambda_s);
                f32x2 ss1_2 = pack2(0.f, 0.f);
                float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
                for (int j = 0; j < V; j++) {
                    const float es = (j == 0 && h0) ? eta_s0 : eta_s1;  // only chunk 0 of lanes 0,1 is in the first half
                    const f32x2 nes2 = pack2(-es, -es);
                    f32x2 ssj = pack2(0.f, 0.f), qnn[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const f32x2 gs = fma2(ne2, pp[j][h], mul2(ls2, qq[j][h]));
                        if (j == 0)
                            ssj = fma2(gs, gs, ssj);
                        else
                            ss1_2 = fma2(gs, gs, ss1_2);
                        qnn[h] = fma2(nes2, gs, qq[j][h]);
                    }
                    if (ready && act[j] && (full || (j == 0 && h0))) {
                        float4 v;
                        unpack2(qnn[0], v.x, v.y);
                        unpack2(qnn[1], v.z, v.w);
                        srow[l + L * j] = v;
                    }
                    if (j == 0) {
                        const float ss = sum2(ssj);
                        if (h0)
                            ss0 = ss;
                        else
                            ss1 = ss;
                    }
                }
                ss1 += sum2(ss1_2);
                ss0 += __shfl_xor_sync(kFullMask, ss0, 1);
                sg.x += ss0 * 0.125f;
                if (full) {
#pragma unroll
                    for (int o = L / 2; o > 0; o >>= 1) ss1 += __shfl_xor_sync(kFullMask, ss1, o);
                    sg.y += ss1 * 0.125f;  // rk_slow for both halves: SURVEY.md F2
                }
                if (ready && leader) s_g[bl] = sg;
                __syncwarp();  // the group's shared-memory stores are ordered before the release of the row
                if (ready && leader)
                    st_release_cta_smem(&s_cnt[bl], DYN ? 0u : ((x1 >> MFK_W1_BBITS) + 1u) & MFK_TICKET_MASK);
            }
            // T side: the row belongs to this group for the whole step; the new row stays in registers for the rest of
            // the run and is written through (the next group to own it may run on another SM)
            {
                const float eta_t0 = g.eta * rsqrtf(tg.x), eta_t1 = g.eta * rsqrtf(tg.y);
                const f32x2 lt2 = pack2(g.lambda_t, g.lambda_t);
                f32x2 st1_2 = pack2(0.f, 0.f);
                float st0 = 0.f, st1 = 0.f;
                float4 *trow = reinterpret_cast<float4 *>(g.T + (size_t)cur_row * k_al);
#pragma unroll
                for (int j = 0; j < V; j++) {
                    const float et = (j == 0 && h0) ? eta_t0 : eta_t1;
                    const f32x2 net2 = pack2(-et, -et);
                    f32x2 stj = pack2(0.f, 0.f), pnn[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const f32x2 gt = fma2(ne2, qq[j][h], mul2(lt2, pp[j][h]));
                        if (j == 0)
                            stj = fma2(gt, gt, stj);
                        else
                            st1_2 = fma2(gt, gt, st1_2);
                        pnn[h] = fma2(net2, gt, pp[j][h]);
                    }
                    if (ready && act[j] && (full || (j == 0 && h0))) {
                        float4 v;
                        unpack2(pnn[0], v.x, v.y);
                        unpack2(pnn[1], v.z, v.w);
                        p[j] = v;
                        __stcg(trow + l + L * j, v);
                    }
                    if (j == 0) {
                        const float st = sum2(stj);
                        if (h0)
                            st0 = st;
                        else
                            st1 = st;
                    }
                }
                st1 += sum2(st1_2);
                st0 += __shfl_xor_sync(kFullMask, st0, 1);
                float2 tgn = make_float2(tg.x + st0 * 0.125f, tg.y);
                if (full) {
#pragma unroll
                    for (int o = L / 2; o > 0; o >>= 1) st1 += __shfl_xor_sync(kFullMask, st1, o);
                    tgn.y += st1 * 0.125f;
                }
                if (ready) {
                    tg = tgn;
                    if (leader) __stcg(reinterpret_cast<float2 *>(g.TG) + cur_row, tgn);
                    done |= 1u << sel;
                }
            }
        }

        // ---- stage the S band out ----
        if (dead) s_dead = 1;
        __syncthreads();
        {
            float4 *dst = reinterpret_cast<float4 *>(g.S) + (size_t)row0 * nvec;
            for (int i = tid; i < nrows * nvec; i += blockDim.x) __stcg(dst + i, s_rows[i]);
            float2 *dstg = reinterpret_cast<float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) __stcg(dstg + i, s_g[i]);
        }
        const int any_dead = s_dead;
        __syncthreads();
        if (any_dead) break;  // CTA-uniform: no warp goes on to a pass its siblings have left
    }

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
unsigned mfk_sgd_run_slot_bytes(int k_al, int groups) { return (unsigned)groups * (unsigned)(k_al * 4 + 16) + 16u; }

int mfk_sgd_run_supported(int k_al, int L_, int fun, float lambda1_s, float lambda1_t, int do_nmf) {
    return k_al <= 128 && L_ == 8 && fun == MFK_FUN_L2_MFR && lambda1_s == 0.f && lambda1_t == 0.f && !do_nmf;
}

int mfk_sgd_run_epoch(const mfk_band_args *args, void *stream) {
    const bool st = args->stats != nullptr, dy = args->dynamic != 0;
    const void *fn = dy ? (st ? (const void *)k_sgd_run_epoch<true, true> : (const void *)k_sgd_run_epoch<true, false>)
                        : (st ? (const void *)k_sgd_run_epoch<false, true> : (const void *)k_sgd_run_epoch<false, false>);
    if (!mfk_sgd_run_supported(args->k_al, args->shape.L, args->fun, args->lambda1_s, args->lambda1_t, args->do_nmf))
        return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)args->shape.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block(args->shape.nWarps * 32);
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, args->shape.smem_bytes, (cudaStream_t)stream);
}

}  // extern "C"
