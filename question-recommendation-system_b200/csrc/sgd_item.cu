// csrc/sgd_item.cu -- the throughput SGD kernel for launches with FEW S ROWS ("item" kernel), sm_100a.
//
// Same job as k_sgd_run_epoch (csrc/sgd_run.cu: SolverBase::run + L2_MFR::prepare_for_sg_update + MFSolver::sg_update,
// mf/mf.cpp:1220-1235, 1720-1728, 1462-1548, and the scheduler's guarantee that no two concurrent updates share a row or a
// column, mf/mf.cpp:130-142), for the default loss (L2_MFR, no L1 term, no NMF) and k_al <= 128 -- organised around the
// side that has FEW rows.
//
// Why.  When the item stripes of config #3 rotate over 8 GPUs a launch sees 2 225 item rows and 60 000 user rows for 1.56M
// ratings: 702 ratings per item row, and a conflict-free schedule performs them one after the other whatever else it does.
// The run kernel hands an item row from group to group through a lock in shared memory; measured, a row changes hands once
// per 1.35 us (an iteration of the waiting group) although the update itself holds it ~0.4 us -- 0.95 ms per launch, 29 %
// of the group-iterations doing work.  Here an S row is never handed over at all:
//
//   * every S row belongs to ONE group for the whole launch (group g of CTA c owns the rows r = g mod nG of band c).  The
//     group walks all ratings of the row back to back with the row and its accumulators in REGISTERS: no S-row lock, no
//     shared-memory copy of the S band, no limit on the band size (config #4's ten passes per launch are gone), and the
//     time of a launch is (ratings per S row) x (one dependent update), not x (one iteration of somebody else).
//   * T rows are taken one by one through lock words in global memory (the T-row locks of sgd_run.cu): the lock of an
//     entry is asked for two entries ahead (relaxed CAS at L2, answer looked at an iteration later), the locked row travels
//     to one of the group's two slots in shared memory by cp.async while the entry before it is computed, and the lock
//     goes back through the CTA's releaser warp (one fence.acq_rel.gpu per batch of returned rows, none in the working
//     warps).  Locks are asked for in stream order and a group that waits holds only rows of earlier entries: no cycle.
//   * the stream of a group is ordered by (S row, rotated T row): the walk over the T rows of an S row starts at a place
//     that depends on the row, so the groups do not all start with the same few T rows.
//
// Arithmetic per rating exactly as in the run kernel: z by butterfly shuffle inside the group,
// e = r - z, both gradients from the OLD p and q, G += sum(g^2)/8 for BOTH halves, dims 0-7 and 8..k_al with
// separate accumulators, epoch 0 touches dims 0-7 only, loss in double.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_helpers.cuh"
#include "kernels.h"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr unsigned kNoRow = 0xffffffffu;
constexpr unsigned kBMask = (1u << MFK_W1_BBITS) - 1u;
constexpr int kRelQ = 128;     // entries of the release queue (a power of two)
constexpr int kAhead = 3;      // entries past the head whose lock may be held or asked for

typedef ulonglong2 chunk_t;

__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// L: lanes per group (8: four groups = four S rows per warp, 16 dimensions per lane at k = 128; 32: the warp is the group,
// 4 dimensions per lane -- a launch with few S rows per CTA has few independent chains of updates, and a chain is short
// when many lanes share its arithmetic and several warps share a scheduler).  NW: working warps the instantiation is
// compiled for.  KFULL: k_al == 128, every lane owns all its chunks.  FULL: every dimension is updated (false only in epoch 0).
template <int L, int NW, bool STATS, bool KFULL, bool FULL>
__global__ void __launch_bounds__((NW + 1) * 32, 1) k_sgd_item_epoch(const __grid_constant__ mfk_band_args g) {
    constexpr int V = 32 / L;  // 16-byte chunks per lane: covers k_al <= 128
    // STATS (MFB200_STATS=1): [0] warp iterations, [1] of them with an update, [2] group updates; group-iterations without
    // one because [3] the stream is finished, [4] the T row is not here yet (lock or copy), [6] CAS that found the row taken
    unsigned long long st_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const mfk_band_shape &sh = g.shape;
    const int k_al = g.k_al, nvec = k_al >> 2;
    chunk_t *s_slots = reinterpret_cast<chunk_t *>(smem_raw);  // [nG][2][nvec + 1]: row + the 16-byte pair with its accumulators
    __shared__ unsigned s_relq[kRelQ], s_qtail, s_qhead, s_wdone;
    __shared__ int s_dead;

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & (L - 1), gi = lane / L;
    const int nG = sh.nG, gamma = warp * (32 / L) + gi;
    const bool leader = l == 0;
    const bool helper = warp == sh.nWarps;  // the releaser warp
    chunk_t *slot0 = s_slots + (size_t)(gamma < nG ? gamma : 0) * 2 * (nvec + 1);

    bool act[V];
#pragma unroll
    for (int j = 0; j < V; j++) act[j] = KFULL || l + L * j < nvec;
    const bool h0 = l < 2;        // chunk 0 of lanes 0,1 = dims 0-7: the first AdaGrad half
    const bool st0 = FULL || h0;  // chunk 0 is stored

    const float eta = g.eta;
    float *const Tbase = g.T;
    float2 *const TGbase = reinterpret_cast<float2 *>(g.TG);
    double loss = 0.0;
    float lossf = 0.f;
    if (tid == 0) {
        s_dead = 0;
        s_qtail = 0u;
        s_qhead = 0u;
        s_wdone = 0u;
    }
    for (int i = tid; i < kRelQ; i += blockDim.x) s_relq[i] = 0u;
    __syncthreads();

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
        const int sb = c;  // one pass: S band c
        const int row0 = sb * sh.segS;
        const unsigned pos = g.goff[(size_t)sb * nG + gamma];
        const unsigned end = g.goff[(size_t)sb * nG + gamma + 1];

        // the stream window: two batches of L entries in registers (lane l holds entry l of each batch)
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
        auto pf_rows = [&](unsigned bbase, unsigned w0) {  // pull the T rows of a batch into L2
            if (bbase + (unsigned)l < end) {
                const char *rp = reinterpret_cast<const char *>(Tbase + (size_t)w0 * k_al);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (KFULL || i * 128 < k_al * 4) prefetch_l2(rp + i * 128);
                prefetch_l2(TGbase + w0);
            }
        };
        unsigned cbase = pos;
        ld_batch(cbase, x0, x1, xr);
        ld_batch(cbase + L, y0, y1, yr);
        pf_rows(cbase, x0);
        pf_rows(cbase + L, y0);
        unsigned nb = cbase < end ? min((unsigned)L, end - cbase) : 0u;  // entries in the current batch
        int hs = 0;  // head: first entry of the window that has not been computed
        int cp = 0;  // first entry whose row has not been sent to a slot (hs <= cp <= lk, cp <= hs + 2)
        int lk = 0;  // first entry no lock covers yet (lk <= hs + kAhead)
        unsigned cas_old = 1u, cas_row = kNoRow;  // the CAS of the previous iteration: its row (kNoRow: none), the leader's answer
        unsigned last_lk = kNoRow;    // T row of entry lk - 1 (a repeated T row: one lock covers both entries)
        unsigned parity = 0u;         // slot of the head entry
        // the S row in registers
        unsigned cur_s = kNoRow;
        chunk_t q[V];
        float2 sg = make_float2(1.f, 1.f);
#pragma unroll
        for (int j = 0; j < V; j++) q[j] = make_ulonglong2(0ull, 0ull);
        unsigned idle = 0;
        unsigned long long idle_since = 0;
        bool dead = false;

        auto flush_q = [&]() {  // the S row goes home (the group is its only writer in this launch)
            if (cur_s != kNoRow) {
                chunk_t *srow = reinterpret_cast<chunk_t *>(g.S + (size_t)(row0 + (int)cur_s) * k_al);
#pragma unroll
                for (int j = 0; j < V; j++)
                    if (act[j] && (j == 0 ? st0 : FULL)) __stcg(srow + l + L * j, q[j]);
                if (leader) __stcg(reinterpret_cast<float2 *>(g.SG) + row0 + (int)cur_s, sg);
            }
        };

        for (;;) {
            // (1) current batch used up: the next one becomes current and a new next one is requested
            if (hs == (int)nb && nb != 0u) {
                cbase += L;
                hs = 0;
                cp = max(cp - L, 0);
                lk = max(lk - L, 0);
                x0 = y0; x1 = y1; xr = yr;
                ld_batch(cbase + L, y0, y1, yr);
                pf_rows(cbase + L, y0);
                nb = cbase < end ? min((unsigned)L, end - cbase) : 0u;
                loss += (double)lossf;
                lossf = 0.f;
            }
            const bool valid = hs < (int)nb;
            const unsigned hw0 = __shfl_sync(kFullMask, x0, hs & (L - 1), L);
            const unsigned hw1 = __shfl_sync(kFullMask, x1, hs & (L - 1), L);
            const float r = __shfl_sync(kFullMask, xr, hs & (L - 1), L);

            // (2) the lock asked for one iteration ago
            {
                const unsigned old = __shfl_sync(kFullMask, cas_old, 0, L);
                if (cas_row != kNoRow) {
                    if (old == 0u) {
                        lk++;
                        last_lk = cas_row;
                    } else if (STATS && leader) {
                        st_[6]++;
                    }
                    cas_row = kNoRow;
                }
            }
            // the head can be computed when its row was sent to its slot in an EARLIER iteration (wait_group 1 below leaves
            // only this iteration's copy in flight)
            const bool ready = valid && cp > hs;
            // (3) the row of the first locked entry that is not on its way starts for its slot (slot = entry parity).  An
            // entry with the head's T row (a repeated row: one lock covers both) waits until the head's stores are out.
            {
                const unsigned cw0 = __shfl_sync(kFullMask, cp < L ? x0 : y0, cp & (L - 1), L);
                if (cp < lk && cp < hs + 2 && !(cp > hs && cw0 == hw0)) {
                    chunk_t *slot = slot0 + (size_t)((parity + (unsigned)(cp - hs)) & 1u) * (nvec + 1);
                    const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)cw0 * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) cp_async16(slot + l + L * j, trow + l + L * j);
                    if (leader)
                        cp_async16(slot + nvec, reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(TGbase + cw0) & ~(uintptr_t)15));
                    cp++;
                }
                cp_async_commit();  // one group per iteration, empty or not: "all but the newest" = everything older
            }
            // (4) the next lock: same T row as the entry before it (a duplicate rating), or a CAS for its row
            {
                const unsigned lw0 = __shfl_sync(kFullMask, lk < L ? x0 : y0, lk & (L - 1), L);
                if (lk < hs + kAhead && lk < 2 * L && cbase + (unsigned)lk < end) {
                    if (lw0 == last_lk) {
                        lk++;
                    } else {
                        cas_row = lw0;
                        if (leader) cas_old = cas_relaxed_gpu(g.tlock + lw0, 0u, 1u);
                    }
                }
            }
            if (STATS) {
                if (lane == 0) st_[0]++;
                if (leader) {
                    if (ready) st_[2]++;
                    else if (!valid) st_[3]++;
                    else st_[4]++;
                }
            }
            if (!__any_sync(kFullMask, ready)) {
                if (__all_sync(kFullMask, nb == 0u)) break;
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

            // (5) the head's T row: out of its slot (the copies of earlier iterations are complete)
            cp_async_wait_1();
            __syncwarp();
            chunk_t p[V];
            float2 tg;
            {
                const chunk_t *slot = slot0 + (size_t)(parity & 1u) * (nvec + 1);
#pragma unroll
                for (int j = 0; j < V; j++) p[j] = act[j] ? slot[l + L * j] : make_ulonglong2(0ull, 0ull);
                const float4 pair = *reinterpret_cast<const float4 *>(slot + nvec);
                const bool odd = ((reinterpret_cast<uintptr_t>(TGbase + hw0) >> 3) & 1u) != 0;
                tg = odd ? make_float2(pair.z, pair.w) : make_float2(pair.x, pair.y);
            }
            // (6) the head's S row: in registers for all its ratings
            const unsigned bl = hw1 & kBMask;
            if (ready && bl != cur_s) {
                flush_q();
                const chunk_t *srow = reinterpret_cast<const chunk_t *>(g.S + (size_t)(row0 + (int)bl) * k_al);
#pragma unroll
                for (int j = 0; j < V; j++) q[j] = act[j] ? __ldcg(srow + l + L * j) : make_ulonglong2(0ull, 0ull);
                sg = __ldcg(reinterpret_cast<const float2 *>(g.SG) + row0 + (int)bl);
                cur_s = bl;
            }

            // ---- the update (sg_update, mf/mf.cpp:1462-1548, 1228-1234); a group that is not ready computes on whatever
            // its registers hold and stores nothing ----
            float pp_all, pp0, pq_all, pq0, qq_all, qq0;
            {
                f32x2 na = mul2(p[0].x, p[0].x), da = mul2(p[0].x, q[0].x), qa = mul2(q[0].x, q[0].x);
                na = fma2(p[0].y, p[0].y, na);
                da = fma2(p[0].y, q[0].y, da);
                qa = fma2(q[0].y, q[0].y, qa);
                f32x2 nbv = pack2(0.f, 0.f), db = nbv, qb = nbv;
#pragma unroll
                for (int j = 1; j < V; j++) {
                    nbv = fma2(p[j].x, p[j].x, nbv);
                    db = fma2(p[j].x, q[j].x, db);
                    qb = fma2(q[j].x, q[j].x, qb);
                    nbv = fma2(p[j].y, p[j].y, nbv);
                    db = fma2(p[j].y, q[j].y, db);
                    qb = fma2(q[j].y, q[j].y, qb);
                }
                const float c0n = sum2(na), d0 = sum2(da), q0 = sum2(qa);
                pp0 = h0 ? c0n : 0.f;
                pq0 = h0 ? d0 : 0.f;
                qq0 = h0 ? q0 : 0.f;
                pp_all = c0n + sum2(nbv);
                pq_all = d0 + sum2(db);
                qq_all = q0 + sum2(qb);
                pp0 += __shfl_xor_sync(kFullMask, pp0, 1);  // (dims 0-7 = chunk 0 of lanes 0,1 for either L)
                pq0 += __shfl_xor_sync(kFullMask, pq0, 1);
                qq0 += __shfl_xor_sync(kFullMask, qq0, 1);
#pragma unroll
                for (int o = L / 2; o > 0; o >>= 1) {
                    pp_all += __shfl_xor_sync(kFullMask, pp_all, o);
                    pq_all += __shfl_xor_sync(kFullMask, pq_all, o);
                    qq_all += __shfl_xor_sync(kFullMask, qq_all, o);
                }
            }
            const float e = r - pq_all;  // mf/mf.cpp:1724 (z = <p,q>, calc_z 1264-1273)
            // (the S row's accumulators are kept up to date by the leader; every lane read the T row's pair from the slot)
            const float sgx = __shfl_sync(kFullMask, sg.x, 0, L), sgy = __shfl_sync(kFullMask, sg.y, 0, L);
            if (ready) {
                lossf = fmaf(e, e, lossf);
                const float et0 = eta * rsqrtf(tg.x), et1 = FULL ? eta * rsqrtf(tg.y) : 0.f;
                const float es0 = eta * rsqrtf(sgx), es1 = FULL ? eta * rsqrtf(sgy) : 0.f;
                // q' = (1 - eta_s lambda_s) q + (eta_s e) p ; p' = (1 - eta_t lambda_t) p + (eta_t e) q, both from the OLD rows
                const float esa = h0 ? es0 : es1, eta_a = h0 ? et0 : et1;
                const float ks1 = fmaf(-es1, g.lambda_s, 1.f), ks2 = es1 * e, ksa1 = fmaf(-esa, g.lambda_s, 1.f), ksa2 = esa * e;
                const float kt1 = fmaf(-et1, g.lambda_t, 1.f), kt2 = et1 * e, kta1 = fmaf(-eta_a, g.lambda_t, 1.f), kta2 = eta_a * e;
                const f32x2 ks1v = pack2(ks1, ks1), ks2v = pack2(ks2, ks2), ksa1v = pack2(ksa1, ksa1), ksa2v = pack2(ksa2, ksa2);
                const f32x2 kt1v = pack2(kt1, kt1), kt2v = pack2(kt2, kt2), kta1v = pack2(kta1, kta1), kta2v = pack2(kta2, kta2);
                chunk_t *trow = reinterpret_cast<chunk_t *>(Tbase + (size_t)hw0 * k_al);
#pragma unroll
                for (int j = 0; j < V; j++) {
                    chunk_t qn, pn;
                    qn.x = fma2(j == 0 ? ksa2v : ks2v, p[j].x, mul2(j == 0 ? ksa1v : ks1v, q[j].x));
                    qn.y = fma2(j == 0 ? ksa2v : ks2v, p[j].y, mul2(j == 0 ? ksa1v : ks1v, q[j].y));
                    pn.x = fma2(j == 0 ? kta2v : kt2v, q[j].x, mul2(j == 0 ? kta1v : kt1v, p[j].x));
                    pn.y = fma2(j == 0 ? kta2v : kt2v, q[j].y, mul2(j == 0 ? kta1v : kt1v, p[j].y));
                    if (act[j] && (j == 0 ? st0 : FULL)) {
                        __stcg(trow + l + L * j, pn);
                        q[j] = qn;
                    }
                }
                if (leader) {
                    // G += sum(g^2) / 8 for BOTH halves (the shipped SSE path's rk);
                    // sum g_s^2 = lambda_s^2 <q,q> - 2 lambda_s e <p,q> + e^2 <p,p>
                    const float e2 = e * e;
                    {
                        const float ls = g.lambda_s, m2 = -2.f * ls * e, l2 = ls * ls;
                        sg.x += fmaf(l2, qq0, fmaf(m2, pq0, e2 * pp0)) * 0.125f;
                        if (FULL) sg.y += fmaf(l2, qq_all - qq0, fmaf(m2, pq_all - pq0, e2 * (pp_all - pp0))) * 0.125f;
                    }
                    {
                        const float lt = g.lambda_t, m2 = -2.f * lt * e, l2 = lt * lt;
                        tg.x += fmaf(l2, pp0, fmaf(m2, pq0, e2 * qq0)) * 0.125f;
                        if (FULL) tg.y += fmaf(l2, pp_all - pp0, fmaf(m2, pq_all - pq0, e2 * (qq_all - qq0))) * 0.125f;
                        __stcg(TGbase + hw0, tg);
                    }
                }
            }
            // (7) the T row goes back unless the next entry is the same row again (a duplicate rating: its row is then read
            // from global memory by the copy that has already been issued -- the stores above must be visible to it, so a
            // duplicate waits for them: handled by keeping duplicates out of the copy stage until the head has passed)
            {
                const int nidx = hs + 1;
                const unsigned nw0 = __shfl_sync(kFullMask, nidx < L ? x0 : y0, nidx & (L - 1), L);
                const bool same_next = cbase + (unsigned)nidx < end && nw0 == hw0;
                if (__any_sync(kFullMask, ready && !same_next)) {
                    __syncwarp();  // every lane's row stores are ordered before the leader's queue entry
                    if (ready && !same_next && leader) {
                        const unsigned t = atomicAdd(&s_qtail, 1u);
                        while ((int)(t - *reinterpret_cast<volatile unsigned *>(&s_qhead)) >= kRelQ) {}
                        st_release_cta_smem(&s_relq[t & (kRelQ - 1)], hw0 + 1u);
                    }
                }
            }
            if (ready) {
                hs++;
                parity ^= 1u;
            }
        }
        loss += (double)lossf;
        lossf = 0.f;
        flush_q();
        if (dead) s_dead = 1;
        __syncwarp();
        if (lane == 0) atomicAdd(&s_wdone, 1u);
    }
    __syncthreads();

    if (!leader || helper) loss = 0.0;  // every lane of a group accumulated the group's e*e
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

unsigned mfk_sgd_item_smem_bytes(int k_al, int groups) { return (unsigned)groups * 2u * (unsigned)(k_al * 4 + 16); }

int mfk_sgd_item_max_warps(int lanes) { return lanes == 32 ? 24 : 16; }

int mfk_sgd_item_epoch(const mfk_band_args *args, void *stream) {
    const bool st = args->stats != nullptr, kf = args->k_al == 128, fu = args->full != 0, wide = args->shape.L == 32;
    if (!mfk_sgd_run_supported(args->k_al, 8, args->fun, args->lambda1_s, args->lambda1_t, args->do_nmf) ||
        (args->shape.L != 8 && args->shape.L != 32) || !args->dynamic || !args->tlock || args->shape.by_row != 4 ||
        args->shape.nWarps > mfk_sgd_item_max_warps(args->shape.L) || args->shape.nPass != 1)
        return (int)cudaErrorInvalidValue;
#define MFB_ITEM3(LL, WW, S, K) (fu ? (const void *)k_sgd_item_epoch<LL, WW, S, K, true> : (const void *)k_sgd_item_epoch<LL, WW, S, K, false>)
#define MFB_ITEM2(S, K) (wide ? MFB_ITEM3(32, 24, S, K) : MFB_ITEM3(8, 16, S, K))
    const void *fn = st ? (kf ? MFB_ITEM2(true, true) : MFB_ITEM2(true, false)) : (kf ? MFB_ITEM2(false, true) : MFB_ITEM2(false, false));
#undef MFB_ITEM2
#undef MFB_ITEM3
    const unsigned smem = mfk_sgd_item_smem_bytes(args->k_al, args->shape.nG);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block((args->shape.nWarps + 1) * 32);  // (+ the releaser warp)
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, smem, (cudaStream_t)stream);
}

}  // extern "C"
