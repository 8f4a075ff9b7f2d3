// csrc/sgd_cell.cu -- the throughput SGD kernel for SMALL launches ("cell" kernel), sm_100a.
//
// Same job as k_sgd_run_epoch (sgd_run.cu): the per-rating loop SolverBase::run + L2_MFR::prepare_for_sg_update +
// MFSolver::sg_update (mf/mf.cpp:1220-1235, 1720-1728, 1462-1548) under a schedule in which no two concurrent updates
// share a row or a column (what the reference's block scheduler guarantees, mf/mf.cpp:130-142), for the default loss
// (L2_MFR, no L1 term, no NMF) at k_al <= 128, S rows handed out by locks.
//
// Why another kernel.  When the item stripes rotate over several GPUs a launch trains a (users / G) x (items / G) block:
// 1.56M ratings at the Netflix shape on 8 GPUs.  In the band / run kernels the unit that owns a T sub-band for a step is
// a GROUP (8 lanes): nC x nTB x 64 cells of ~4 ratings each, every group walks every step (at least one warp iteration
// per step even when its cell is empty), and cell sizes are Poisson, so a group spends more iterations waiting for the
// sub-band than updating (MFB200_STATS: 37 % "band not released", profiles/r2_run_vs_band_shapes.txt).  Here the unit
// is the CTA:
//
//   * CELLS.  The stream of an S band is ordered by (step, T row): all ratings that CTA c trains while it owns T band
//     (c*S1 + t) mod nTB are one contiguous cell.  The 64 groups of the CTA take CHUNKS of C consecutive entries off a
//     cursor in shared memory (one atomic per chunk, claimed one chunk ahead), so the work of a cell is shared
//     dynamically and a group never walks an empty step.
//   * T-row exclusivity inside the CTA comes from the order: the ratings of one T row inside a cell are adjacent (a
//     run); the group whose chunk holds the run's first entry processes the whole run, also past the end of its chunk,
//     and the next chunk's owner skips the entries that belong to a run begun earlier (bit 31 of w0 marks a run's first
//     entry, bit 30 says that the run goes on after this entry).  The T row stays in registers over the run.
//   * HAND-OFF PER CTA.  s_rem[t] counts the entries of cell t that are not yet done; a group subtracts what it has
//     finished (after a gpu-scope fence that orders its T-row stores) whenever it moves on to another step or has to
//     wait.  The CTA's flag = number of leading steps whose cell is complete AND whose own dependency has been verified
//     (the rule recorded in DESIGN.md section 4, "a bug worth recording"), raised with a max-reduction by whichever
//     group observes it.  The waiting side is as in the run kernel: a predicated poll one iteration ahead of its use,
//     one fence per warp iteration; what a group learns it shares through s_tok.
//   * Because no group walks steps, many fine steps are cheap: S1 (T bands per CTA) is 8 here.  A ring of nC CTAs cannot
//     finish a launch faster than nC * S1 / (S1 - 1) hand-off latencies (each T band must visit every CTA in turn), so
//     the host picks fewer CTAs for small launches (plan_band).
//
// Arithmetic per rating exactly as in sgd_run.cu: z by butterfly shuffle inside the group,
// e = r - z, both gradients from the OLD p and q, G += sum(g^2)/8 for BOTH halves, epoch 0 touches dims
// 0-7 only, loss in double.
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
constexpr unsigned kRowMask = MFK_CELL_ROW_MASK;

typedef ulonglong2 chunk_t;

__device__ __forceinline__ unsigned ld_volatile_smem(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void red_max_relaxed_gpu(unsigned *p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// control words in shared memory
enum { CTL_CUR = 0, CTL_PUB = 1, CTL_TOK = 2, CTL_DEAD = 3 };

template <bool STATS, bool KFULL, bool FULL>
__global__ void __launch_bounds__(512, 1) k_sgd_cell_epoch(const __grid_constant__ mfk_band_args g) {
    // STATS (MFB200_STATS=1): [0] warp iterations, [1] of them with an update, [2] group updates; group-iterations
    // without one because [3] the stream is finished, [4] the T band is not released yet, [5] the S row is busy;
    // [6] runs started from the prefetch slot, [7] runs started with a direct (exposed) load.
    // [8] clock cycles of warp iterations with an update, [9] of those without, [10] flag acquisitions, [11] flag raises
    unsigned long long st_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long st_clk = 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const mfk_band_shape &sh = g.shape;
    const int k_al = g.k_al, nvec = k_al >> 2;
    chunk_t *s_rows = reinterpret_cast<chunk_t *>(smem_raw);                        // [rows_cap + 1][nvec]
    float2 *s_g = reinterpret_cast<float2 *>(s_rows + (sh.rows_cap + 1) * nvec);    // [rows_cap + 1]
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_g + sh.rows_cap + 1);          // [rows_cap + 1]
    chunk_t *s_slots = reinterpret_cast<chunk_t *>(smem_raw + ((((size_t)(sh.rows_cap + 1) * (nvec * 16 + 12)) + 15) & ~(size_t)15));
    const int nG = sh.nG;
    unsigned *s_rem = reinterpret_cast<unsigned *>(s_slots + (size_t)nG * (nvec + 1));  // [nTB] entries of a cell not yet done
    unsigned *s_ctl = s_rem + sh.nTB;                                                    // [4]
    const unsigned dummy = (unsigned)sh.rows_cap;  // zero row with accumulators 1: what a group sitting out computes against

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & (L - 1), gi = lane >> 3;
    const int gamma = warp * 4 + gi;
    const bool leader = l == 0;
    chunk_t *slot = s_slots + gamma * (nvec + 1);

    bool act[V];
#pragma unroll
    for (int j = 0; j < V; j++) act[j] = KFULL || l + L * j < nvec;
    const bool h0 = l < 2;
    const bool st0 = FULL || h0;

    const int nTB = sh.nTB, S1 = sh.S1;
    const unsigned C = (unsigned)sh.chunk;
    unsigned *my_flag = g.flags + c;
    const unsigned *nb_flag = g.flags + (c + 1) % sh.nC;
    const bool ring = sh.nC > 1;
    const float eta = g.eta;
    float *const Tbase = g.T;
    float2 *const TGbase = reinterpret_cast<float2 *>(g.TG);
    double loss = 0.0;
    float lossf = 0.f;
    for (int i = tid; i < nvec; i += blockDim.x) s_rows[dummy * nvec + i] = make_ulonglong2(0ull, 0ull);
    if (tid == 0) {
        s_g[dummy] = make_float2(1.f, 1.f);
        s_cnt[dummy] = 0u;
        s_ctl[CTL_DEAD] = 0u;
    }

    for (int pass = 0; pass < sh.nPass; ++pass) {
        const int sb = pass * sh.nC + c;
        const int row0 = sb * sh.segS;
        const int nrows = max(0, min(sh.segS, g.nS - row0));
        const unsigned *coff = g.goff + (size_t)sb * nTB;  // first entry of every cell of this S band
        const unsigned pos = coff[0], end = coff[nTB];
        const unsigned base = g.base + (unsigned)pass * (unsigned)nTB;
        const int t0 = !ring ? nTB : (pass == 0 ? S1 - 1 : -1);  // steps released before anything has been polled
        // ---- stage the S band in, set the cell counters up ----
        {
            const float4 *src = reinterpret_cast<const float4 *>(g.S) + (size_t)row0 * nvec;
            float4 *dst = reinterpret_cast<float4 *>(s_rows);
            for (int i = tid; i < nrows * nvec; i += blockDim.x) dst[i] = __ldcg(src + i);
            const float2 *srcg = reinterpret_cast<const float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) {
                s_g[i] = __ldcg(srcg + i);
                s_cnt[i] = 0u;
            }
            for (int i = tid; i < nTB; i += blockDim.x) s_rem[i] = coff[i + 1] - coff[i];
            if (tid == 0) {
                s_ctl[CTL_CUR] = pos;
                s_ctl[CTL_PUB] = 0u;
                s_ctl[CTL_TOK] = (unsigned)(t0 + 1);  // stored + 1 so that "nothing released" (-1) is 0
            }
        }
        __syncthreads();

        // the chunk being processed (lane l holds its entry l) and the chunk claimed after it
        unsigned x0 = 0u, x1 = 0u, y0 = 0u, y1 = 0u;
        float xr = 0.f, yr = 0.f;
        unsigned cb = 0u, nb = 0u, hs = 0u;  // base of the current chunk, its number of entries, the entry at work
        unsigned yb = 0u, ynb = 0u, yhs = 0u;  // the next chunk: base, entries, its first run head
        bool fin = false;                    // the cursor has passed the end of the stream: nothing left to claim
        bool started = false;
        // a run that goes on past the end of its chunk: its entries are fetched one at a time
        bool in_tail = false;
        unsigned tz0 = 0u, tz1 = 0u, tpos = 0u;
        float tzr = 0.f;

        unsigned cur_row = kNoRow, pre_row = kNoRow;
        chunk_t p[V];
        float2 tg = make_float2(1.f, 1.f);
#pragma unroll
        for (int j = 0; j < V; j++) p[j] = make_ulonglong2(0ull, 0ull);
        int t_ok = t0;
        unsigned fval = base;
        bool polled = false;
        unsigned cnt = 0u;      // entries finished since the last subtraction from s_rem
        int cnt_step = 0;       // the step they belong to
        unsigned idle = 0;
        unsigned long long idle_since = 0;
        bool dead = false;

        if (STATS) st_clk = clock64();
        for (;;) {
            // (1) the current chunk is used up: the next one becomes current and another one is claimed.  Two claims at
            // the start.  A chunk begins at its first run head: entries before it belong to a run begun in an earlier
            // chunk and are processed by that chunk's owner.
            const bool want_swap = !in_tail && hs >= nb && !fin;
            // ONE claim per warp serves every group of the warp that wants a chunk, and those groups get adjacent chunks:
            // the groups of a warp share an instruction stream, so little work should sit in few warps -- four entries
            // in one warp cost one warp iteration, in four warps four (measured: 1.75 updates per updating iteration
            // with a claim per group at 18 entries per cell).
            const unsigned wantm = __ballot_sync(kFullMask, want_swap && leader);
            if (wantm != 0u) {
                const unsigned per = started ? C : 2u * C, nwant = (unsigned)__popc(wantm);
                unsigned nbase = 0u;
                if (lane == 0) nbase = atomicAdd(&s_ctl[CTL_CUR], nwant * per);
                nbase = __shfl_sync(kFullMask, nbase, 0);
                const unsigned rank = (unsigned)__popc(wantm & ((1u << (gi * 8)) - 1u));  // wanting groups before this one
                nbase += rank * C;  // (first claim: the warp's first chunks are adjacent, then its second chunks)
                if (want_swap) {
                    if (started) {
                        x0 = y0; x1 = y1; xr = yr;
                        cb = yb; nb = ynb; hs = yhs;
                    } else {  // first claim of the pass: two chunks at once, the first one is loaded here
                        cb = nbase;
                        nb = cb < end ? min(C, end - cb) : 0u;
                        x0 = 0u; x1 = 0u; xr = 0.f;
                        if ((unsigned)l < nb) {
                            x0 = __ldcs(g.w0 + cb + l);
                            x1 = __ldcs(g.w1 + cb + l);
                            xr = __ldcs(g.rr + cb + l);
                        }
                        nbase += nwant * C;
                    }
                    yb = nbase;
                    ynb = yb < end ? min(C, end - yb) : 0u;
                    y0 = 0u; y1 = 0u; yr = 0.f;
                    if ((unsigned)l < ynb) {
                        y0 = __ldcs(g.w0 + yb + l);
                        y1 = __ldcs(g.w1 + yb + l);
                        yr = __ldcs(g.rr + yb + l);
                    }
                    if (nb == 0u) fin = true;  // the cursor only grows: every later claim would be empty too
                    loss += (double)lossf;
                    lossf = 0.f;
                }
                if (!started) {  // (warp-uniform: all groups start together) first run head of the first chunk
                    const unsigned hb = (__ballot_sync(kFullMask, (unsigned)l < nb && (x0 >> 31) != 0u) >> (gi * 8)) & 0xffu;
                    hs = hb ? (unsigned)__ffs((int)hb) - 1u : nb;
                    started = true;
                }
            }

            // (2) the entry at work and the one after it
            const bool valid = in_tail || hs < nb;
            const unsigned sx0 = __shfl_sync(kFullMask, x0, hs & (L - 1), L);
            const unsigned sx1 = __shfl_sync(kFullMask, x1, hs & (L - 1), L);
            const float sxr = __shfl_sync(kFullMask, xr, hs & (L - 1), L);
            const unsigned hw0 = in_tail ? tz0 : sx0, hw1 = in_tail ? tz1 : sx1;
            const float r = in_tail ? tzr : sxr;
            const int ht = valid ? (int)(hw1 >> MFK_W1_BBITS) : nTB;
            const unsigned hrow = hw0 & kRowMask;
            const bool hcont = valid && (hw0 & MFK_CELL_CONT) != 0u;  // the run goes on after this entry
            const unsigned nidx = hs + 1u;
            const bool nin = !in_tail && nidx < nb;

            // (3) hand-off between CTAs.  Acquiring side: what the other groups of the CTA have learnt (s_tok) and the
            // neighbour's flag as polled one iteration ago.
            bool acquired = false;
            {
                const int tk = (int)ld_acquire_cta_smem(&s_ctl[CTL_TOK]) - 1;
                if (tk > t_ok) t_ok = tk;
                if (polled) {
                    const int s_rel = (int)(fval - base) + S1 - 1;
                    if (s_rel > t_ok) {
                        t_ok = s_rel;
                        acquired = true;
                    }
                }
            }
            const bool can = valid && ht <= t_ok;
            if (__any_sync(kFullMask, acquired)) {
                // acquire side of the flag consumed above: this lane's T-row loads below are ordered after it
                fence_acq_rel_gpu();
                if (acquired && leader) atomicMax(&s_ctl[CTL_TOK], (unsigned)(t_ok + 1));
                if (STATS && acquired && leader) st_[10]++;
            }

            // (4) a new run: its T row comes from the prefetch slot or straight from L2
            const bool sw = can && hrow != cur_row;
            if (__any_sync(kFullMask, sw)) {
                cp_async_wait_all();
                __syncwarp();
                if (sw) {
                    if (pre_row == hrow) {
#pragma unroll
                        for (int j = 0; j < V; j++)
                            if (act[j]) p[j] = slot[l + L * j];
                        const float4 pair = *reinterpret_cast<const float4 *>(slot + nvec);
                        const bool odd = ((reinterpret_cast<uintptr_t>(TGbase + hrow) >> 3) & 1u) != 0;
                        tg = odd ? make_float2(pair.z, pair.w) : make_float2(pair.x, pair.y);
                        pre_row = kNoRow;
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
                __syncwarp();
            }

            // (5) the entry after this one: if it starts another run and its step has been released, its T row starts
            // travelling to the slot now; the poll the NEXT iteration consumes; the first run head of the next chunk
            {
                // (the next chunk was requested at the top of this or of an earlier iteration: this is the first look at it)
                const unsigned hb = (__ballot_sync(kFullMask, (unsigned)l < ynb && (y0 >> 31) != 0u) >> (gi * 8)) & 0xffu;
                yhs = hb ? (unsigned)__ffs((int)hb) - 1u : ynb;
                // the entry after this one, if it is known without a load: the next one of this chunk, or -- when the run
                // ends with this chunk -- the first run head of the next chunk
                const unsigned nsel = nin ? nidx : yhs;
                const unsigned nw0 = __shfl_sync(kFullMask, nin ? x0 : y0, nsel & (L - 1), L);
                const unsigned nw1 = __shfl_sync(kFullMask, nin ? x1 : y1, nsel & (L - 1), L);
                const bool nvalid = nin || (valid && !hcont && yhs < ynb);
                const unsigned nrow = nw0 & kRowMask;
                const int nt = nvalid ? (int)(nw1 >> MFK_W1_BBITS) : ht;
                if (nvalid && nt <= t_ok && nrow != hrow && nrow != cur_row && pre_row == kNoRow) {
                    const chunk_t *trow = reinterpret_cast<const chunk_t *>(Tbase + (size_t)nrow * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++)
                        if (act[j]) cp_async16(slot + l + L * j, trow + l + L * j);
                    if (leader)
                        cp_async16(slot + nvec, reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(TGbase + nrow) & ~(uintptr_t)15));
                    cp_async_commit();
                    pre_row = nrow;
                }
                polled = ring && t_ok < min(max(ht, nt), nTB - 1);
                if (polled) fval = ld_relaxed_gpu(nb_flag);
            }

            // (6) what does not need the S row: squared norms of the T row per AdaGrad half, its step sizes
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

            // (3b) releasing side, as late as possible (the T-row stores of the previous iteration have had time to
            // complete): what this group has finished is taken off its cell's counter when the group moves on to another
            // step, has to wait, or has nothing left.  A group without work also helps the flag over cells that are
            // empty -- their dependency must have been verified.
            {
                const bool flush = cnt != 0u && !(can && ht == cnt_step);
                const unsigned pub_seen = ld_volatile_smem(&s_ctl[CTL_PUB]);
                const bool lookpub = (acquired || !can) && pub_seen < (unsigned)nTB && (int)pub_seen <= t_ok &&
                                     ld_volatile_smem(&s_rem[min(pub_seen, (unsigned)nTB - 1u)]) == 0u;
                if (__any_sync(kFullMask, flush || lookpub)) {
                    fence_acq_rel_gpu();  // release side of the T-row stores counted below (every lane fences its own)
                    __syncwarp();
                    if (leader) {
                        bool completed = false;
                        if (flush) completed = atomicSub(&s_rem[cnt_step], cnt) == cnt;
                        if (completed || lookpub) {
                            const unsigned p0 = ld_volatile_smem(&s_ctl[CTL_PUB]);
                            const int tok = max(t_ok, (int)ld_volatile_smem(&s_ctl[CTL_TOK]) - 1);
                            unsigned q = p0;
                            while (q < (unsigned)nTB && (int)q <= tok && ld_volatile_smem(&s_rem[q]) == 0u) q++;
                            if (q > p0) {
                                // the other groups fenced their stores before they touched the counters read above; this
                                // fence orders those reads before the flag
                                fence_acq_rel_gpu();
                                atomicMax(&s_ctl[CTL_PUB], q);
                                red_max_relaxed_gpu(my_flag, base + q);
                                if (STATS) st_[11]++;
                            }
                        }
                    }
                    if (flush) cnt = 0u;
                }
            }

            // (7) the S row: whoever asks first
            unsigned got = 0u;
            if (can && leader) got = cas_acquire_cta_smem(&s_cnt[hw1 & kBMask], 0u, 1u) == 0u;
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
                if (STATS && lane == 0) {
                    const long long now = clock64();
                    st_[9] += (unsigned long long)(now - st_clk);
                    st_clk = now;
                }
                if (__all_sync(kFullMask, fin && !in_tail && cnt == 0u && ld_volatile_smem(&s_ctl[CTL_PUB]) >= (unsigned)nTB)) break;
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
                    // this entry is done; where the group goes next
                    cnt_step = ht;
                    cnt++;
                    const unsigned at = in_tail ? tpos : cb + hs;  // stream position of the entry just finished
                    if (in_tail || nidx >= nb) {
                        // the chunk has no further entry.  A run that goes on belongs to this group: fetch its next entry
                        // (rare: the latency of this load is exposed); otherwise the chunk is used up.
                        if (hcont) {
                            tpos = at + 1u;
                            tz0 = __ldcs(g.w0 + tpos);
                            tz1 = __ldcs(g.w1 + tpos);
                            tzr = __ldcs(g.rr + tpos);
                            in_tail = true;
                        } else {
                            in_tail = false;
                        }
                        hs = nb;
                    } else {
                        hs = nidx;
                    }
                }
            }
            if (STATS && lane == 0) {
                const long long now = clock64();
                st_[8] += (unsigned long long)(now - st_clk);
                st_clk = now;
            }
        }
        loss += (double)lossf;
        lossf = 0.f;

        // ---- stage the S band out ----
        if (dead) s_ctl[CTL_DEAD] = 1u;
        __syncthreads();
        {
            float4 *dst = reinterpret_cast<float4 *>(g.S) + (size_t)row0 * nvec;
            const float4 *src = reinterpret_cast<const float4 *>(s_rows);
            for (int i = tid; i < nrows * nvec; i += blockDim.x) __stcg(dst + i, src[i]);
            float2 *dstg = reinterpret_cast<float2 *>(g.SG) + row0;
            for (int i = tid; i < nrows; i += blockDim.x) __stcg(dstg + i, s_g[i]);
        }
        const unsigned any_dead = s_ctl[CTL_DEAD];
        __syncthreads();
        if (any_dead) break;
    }

    if (!leader) loss = 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(kFullMask, loss, o);
    if (lane == 0 && loss != 0.0) atomicAdd(g.loss, loss);
    if (STATS && g.stats) {
#pragma unroll
        for (int i = 0; i < 12; i++) {
            unsigned long long v = st_[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
            if (lane == 0 && v) atomicAdd(g.stats + i, v);
        }
    }
}

}  // namespace

extern "C" {

// shared memory beside the S band: the prefetch slots and the dummy row (as the run kernel), one counter per step,
// four control words
unsigned mfk_sgd_cell_extra_bytes(int k_al, int groups, int nTB) {
    return mfk_sgd_run_slot_bytes(k_al, groups) + 4u * (unsigned)nTB + 16u;
}

int mfk_sgd_cell_epoch(const mfk_band_args *args, void *stream) {
    const bool st = args->stats != nullptr, kf = args->k_al == 128, fu = args->full != 0;
    if (!mfk_sgd_run_supported(args->k_al, args->shape.L, args->fun, args->lambda1_s, args->lambda1_t, args->do_nmf) ||
        !args->dynamic || args->shape.by_row != 2 || args->shape.chunk < 1 || args->shape.chunk > 8)
        return (int)cudaErrorInvalidValue;
#define MFB_CELL2(S, K) (fu ? (const void *)k_sgd_cell_epoch<S, K, true> : (const void *)k_sgd_cell_epoch<S, K, false>)
    const void *fn = st ? (kf ? MFB_CELL2(true, true) : MFB_CELL2(true, false)) : (kf ? MFB_CELL2(false, true) : MFB_CELL2(false, false));
#undef MFB_CELL2
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)args->shape.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block(args->shape.nWarps * 32);
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, args->shape.smem_bytes, (cudaStream_t)stream);
}

}  // extern "C"
