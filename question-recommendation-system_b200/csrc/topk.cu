// csrc/topk.cu -- batched top-k scoring of P.Q^T for the predict path (BASELINE.json config #5).
//
// The reference has no top-k function; its predict path is mf_predict (mf/mf.cpp:4295-4314), a sequential
// fp32 dot product per (user, item) pair, and SURVEY.md 8c defines top-k on top of it: score = mf_predict,
// order = score descending, item id ascending.  The index lists produced here are bit-exact to that.
//
// The one dense contraction of the engine runs on the tensor cores:
//   prep      P rows of the requested users and all Q rows -> bf16 (round to nearest), fp32 row norms
//   pass A    k_topk_gemm<MODE_MAX>: S = P_bf16 . Q_bf16^T on tcgen05 (TMA-fed, accumulator in TMEM); the epilogue
//             keeps, per user and per 256-item tile, the maximum of a rigorous LOWER bound of the exact score,
//             s - eps_u * |q_v|.  tau_u = the topk-th largest of those maxima is a lower bound of the exact
//             topk-th score (k_topk_tau).
//   pass C    the same GEMM; the epilogue appends every item whose UPPER bound s + eps_u * |q_v| reaches tau_u
//             to the user's candidate list -- a superset of the exact top-k.
//   select    exact fp32 re-score of the candidates in the reference's summation order, bitonic sort by
//             (score desc, id asc), first topk out (k_topk_select).
// Error bound: bf16 rounding is relative 2^-8 per operand, so |s - exact| <= (2^-7 + 2^-16) sum|p_d q_d| plus fp32
// accumulation terms of relative order k 2^-22; eps_u |q_v| = 2^-7 * 1.05 * |p_u| |q_v| covers both (Cauchy-Schwarz).
// Rows never seen in training are NaN (mf/mf.cpp:996-999) and score exactly b: NaN items are kept out of the GEMM
// (zero rows, bound -inf) and the lowest-numbered ones join every candidate list; NaN users get items 0..topk-1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int TK_M = 128;        // users per accumulator (UMMA M)
constexpr int TK_N = 256;        // items per tile  (UMMA N)
constexpr int TK_KATOM = 64;     // bf16 elements per 128-byte swizzle atom
constexpr int TK_STAGES = 2;     // Q tile stages in shared memory
constexpr int TK_EPI_WARPS = 16; // 4 TMEM lane quarters x 2 accumulators x 2 column halves
constexpr int TK_THREADS = (2 + TK_EPI_WARPS) * 32;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-17: epilogue
constexpr float TK_EPS = 1.05f / 128.0f;  // 2^-7 * 1.05
constexpr unsigned kFullMask = 0xffffffffu;

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *tm, void *dst, uint64_t *bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {  // implies tcgen05.fence::before_thread_sync
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// TMEM -> registers, asynchronous: the 32 destination registers are valid only after tmem_ld_wait(r), which
// names them as in/out operands so that no use can be scheduled before the wait.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// 8 columns (a group of 8 items) of the warp's 32 lanes; synchronous
__device__ __forceinline__ void tmem_ld8_sync(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: 8-row groups are 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);   // start address
    d |= (uint64_t)(1024u >> 4) << 32;          // stride byte offset
    d |= 1ull << 46;                            // descriptor version (sm_100)
    d |= 2ull << 61;                            // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N = 256
__device__ __forceinline__ uint32_t umma_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK_N >> 3) << 17) | ((uint32_t)(TK_M >> 4) << 24);
}

enum { MODE_MAX = 0, MODE_CAND = 1 };

struct TopkGemmArgs {
    int n_user_tiles;   // tiles of 128*nh users in the batch
    int nh;             // accumulators (halves of 128 users) per CTA tile: 1 or 2
    int n_item_tiles;   // tiles of 256 items
    int tile_stride;    // MODE_MAX: every tile_stride-th item tile is sampled; MODE_CAND: 1
    int tile0;          // first item tile visited (tiles tile0, tile0 + tile_stride, ... < n_item_tiles)
    int gpt;            // MODE_MAX: maxima per tile: 2 (per half tile = epilogue warp) or 8 (32-column chunks)
    int ub;             // users in the batch, padded to 256
    const float *eps;   // [ub] eps_u (0 for padding / NaN users)
    const float *qg;    // [n_item_tiles*8] max |q_v| of every 32-item group; a group that holds a NaN or padding
                        // item is +inf in MODE_MAX (no lower bound from it); such items count as -inf in MODE_CAND
    const float *ag;    // [n_item_tiles*8] bias bound of every 32-item group: the smallest a_i in MODE_MAX, the largest in
                        // MODE_CAND (all zero without centring)
    float *maxes;       // MODE_MAX: [n_sampled_tiles * gpt][ub]
    const float *tau;   // MODE_CAND: [ub] (+inf: no candidates)
    float4 *grp_sc;     // MODE_CAND: [ub][gmax][2] the 8 bf16-GEMM scores of every 8-item group that may hold a candidate
    int *grp_id;        // MODE_CAND: [ub][gmax] its group number (first item / 8)
    int *grp_cnt;       // MODE_CAND: [2][ub]: per half list (gmax/2 entries each); may exceed: overflow
    int gmax;
};

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // one FMNMX3
// maxima of the four 8-column groups of a 32-column chunk: 4 instructions per group
__device__ __forceinline__ void group_maxima(const uint32_t *vr, float g[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float a = max3(__uint_as_float(vr[8 * i + 0]), __uint_as_float(vr[8 * i + 1]), __uint_as_float(vr[8 * i + 2]));
        const float b = max3(__uint_as_float(vr[8 * i + 3]), __uint_as_float(vr[8 * i + 4]), __uint_as_float(vr[8 * i + 5]));
        g[i] = max3(a, b, fmaxf(__uint_as_float(vr[8 * i + 6]), __uint_as_float(vr[8 * i + 7])));
    }
}

// One CTA = one tile of 128*nh users against a stream of 256-item tiles.  Warp 0 feeds shared memory with TMA, one
// thread of warp 1 issues the MMAs (accumulator h = users [128h, 128h+128) of the tile, 256 TMEM columns each, the
// two accumulators alternate so that one is computed while the other is read), 16 epilogue warps read them: warp
// (quarter q, accumulator h, column half c) owns TMEM lanes 32q..32q+31 = one user per thread for the whole kernel
// and 128 of the 256 columns of every item tile.  The epilogue costs ~0.6 instructions per score: 3-input maxima
// over 8-column groups first, the per-item bound only where a group maximum reaches the user's threshold.
template <int KATOMS, int MODE>
__global__ void __launch_bounds__(TK_THREADS, 1)
k_topk_gemm(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, const TopkGemmArgs a) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024-byte aligned
    constexpr uint32_t A_BYTES = KATOMS * TK_M * 128;          // 16 KB per atom, per accumulator
    constexpr uint32_t B_BYTES = KATOMS * TK_N * 128;          // 32 KB per atom
    uint8_t *sA = smem;
    uint8_t *sB = smem + 2 * A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + TK_STAGES * B_BYTES);
    uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6, *a_full = bars + 8, *a_free = bars + 9;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 8);  // one arrival per epilogue warp of that accumulator
        }
        mbar_init(a_full, 1);
        mbar_init(a_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(s_tmem, 512);  // two accumulators of 256 columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int n_it = (a.n_item_tiles - a.tile0 + a.tile_stride - 1) / a.tile_stride;  // item tiles visited per user tile
    const int nh = a.nh;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it_glob = 0, ut_count = 0;
            for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x, ut_count++) {
                if (ut_count > 0) mbar_wait(a_free, (ut_count - 1) & 1);  // the MMAs of the previous user tile are done
                mbar_expect_tx(a_full, A_BYTES * nh);
                for (int h = 0; h < nh; h++)
                    for (int ka = 0; ka < KATOMS; ka++)
                        tma_load_2d(&tmP, sA + h * A_BYTES + ka * (TK_M * 128), a_full, ka * TK_KATOM, (ut * nh + h) * TK_M);
                for (int i = 0; i < n_it; i++, it_glob++) {
                    const int s = it_glob & 1;
                    mbar_wait(&empty[s], ((it_glob >> 1) & 1) ^ 1);
                    mbar_expect_tx(&full[s], B_BYTES);
                    const int tile = a.tile0 + i * a.tile_stride;
                    for (int ka = 0; ka < KATOMS; ka++)
                        tma_load_2d(&tmQ, sB + s * B_BYTES + ka * (TK_N * 128), &full[s], ka * TK_KATOM, tile * TK_N);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc();
            uint32_t it_glob = 0, ut_count = 0;
            for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x, ut_count++) {
                mbar_wait(a_full, ut_count & 1);
                for (int i = 0; i < n_it; i++, it_glob++) {
                    const int s = it_glob & 1;
                    mbar_wait(&full[s], (it_glob >> 1) & 1);
                    for (int h = 0; h < nh; h++) {
                        mbar_wait(&tempty[h], (it_glob & 1) ^ 1);  // accumulator h has been read out
                        tc_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)h * TK_N;
#pragma unroll
                        for (int ka = 0; ka < KATOMS; ka++) {
                            const uint64_t ad = smem_desc_sw128(smem_u32(sA + h * A_BYTES + ka * (TK_M * 128)));
                            const uint64_t bd = smem_desc_sw128(smem_u32(sB + s * B_BYTES + ka * (TK_N * 128)));
#pragma unroll
                            for (int k4 = 0; k4 < TK_KATOM / 16; k4++)  // UMMA K = 16 bf16 = 32 bytes inside the atom
                                umma_bf16(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, (ka | k4) ? 1u : 0u);
                        }
                        umma_commit(&tfull[h]);  // accumulator h may be read when these MMAs are done
                    }
                    umma_commit(&empty[s]);      // ... and the Q stage may be overwritten
                }
                umma_commit(a_free);
            }
        }
    } else {
        // ===== epilogue =====
        const int q4 = warp & 3, sub = (warp - 2) >> 2, h = sub >> 1, c = sub & 1;
        if (h < nh) {
            constexpr int HC = TK_N / 2;  // columns per warp
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)h * TK_N + (uint32_t)c * HC;
            const float ninf = __int_as_float(0xff800000);
            uint32_t it_glob = 0;
            for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x) {
                const int ub = (ut * nh + h) * TK_M + q4 * 32 + lane;
                const float eps = a.eps[ub];
                float tau_s = 0.f, neg_eps = 0.f;
                int cnt = 0;  // MODE_CAND: groups kept by this thread in its column half
                if (MODE == MODE_CAND) {
                    // exact >= tau implies s >= tau - eps|q_v| >= tau - eps max|q| =: thr (a little lower, for the
                    // rounding of the line that computes it)
                    const float tau = a.tau[ub];
                    tau_s = tau - fabsf(tau) * 1e-6f;
                    neg_eps = -1.001f * eps;
                }
                // the two column halves append to disjoint halves of the user's list (gmax/2 each)
                const int my_gmax = a.gmax / 2;
                const size_t my_base = (size_t)ub * a.gmax + (size_t)c * my_gmax;
                for (int i = 0; i < n_it; i++, it_glob++) {
                    const int tile = a.tile0 + i * a.tile_stride;
                    // group norms of this tile: requested before the wait for the accumulator
                    const float4 qg_cur = __ldg(reinterpret_cast<const float4 *>(a.qg) + (size_t)tile * 2 + c);
                    const float qgv[4] = {qg_cur.x, qg_cur.y, qg_cur.z, qg_cur.w};
                    const float4 ag_cur = __ldg(reinterpret_cast<const float4 *>(a.ag) + (size_t)tile * 2 + c);
                    const float agv[4] = {ag_cur.x, ag_cur.y, ag_cur.z, ag_cur.w};
                    uint32_t va[32], vb[32];
                    mbar_wait(&tfull[h], it_glob & 1);
                    tc_fence_after();
                    tmem_ld32(taddr, va);
                    tmem_ld32(taddr + 32, vb);
                    float mx = ninf;
                    auto process = [&](const uint32_t *vr, int ch) {
                        float g[4];
                        group_maxima(vr, g);
                        if (MODE == MODE_MAX) {
                            // s - eps|q_v| >= s - eps max|q|: a lower bound of the best exact score of the chunk
                            const float m32 = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                            // (+ the smallest item bias of the chunk; a little lower for the rounding of the sum)
                            const float lb = fmaf(-eps, qgv[ch], m32 + agv[ch]) - 4e-7f * (fabsf(m32) + fabsf(agv[ch]));
                            if (a.gpt == 8) a.maxes[((size_t)i * 8 + c * 4 + ch) * a.ub + ub] = lb;
                            else mx = fmaxf(mx, lb);
                        } else {
                            // a group whose maximum reaches thr goes to the user's list as it is (8 scores + its
                            // number); the selection kernel looks at the items.  Rare: ~0.6 % of the groups.
                            // (- the largest item bias of the chunk; a little lower for the rounding of the difference)
                            const float thr = fmaf(neg_eps, qgv[ch], tau_s - agv[ch]) - 4e-7f * (fabsf(tau_s) + fabsf(agv[ch]));
#pragma unroll
                            for (int gi = 0; gi < 4; gi++) {
                                if (g[gi] >= thr) {
                                    if (cnt < my_gmax) {
                                        float4 *dst = a.grp_sc + (my_base + cnt) * 2;
                                        dst[0] = make_float4(__uint_as_float(vr[8 * gi + 0]), __uint_as_float(vr[8 * gi + 1]),
                                                             __uint_as_float(vr[8 * gi + 2]), __uint_as_float(vr[8 * gi + 3]));
                                        dst[1] = make_float4(__uint_as_float(vr[8 * gi + 4]), __uint_as_float(vr[8 * gi + 5]),
                                                             __uint_as_float(vr[8 * gi + 6]), __uint_as_float(vr[8 * gi + 7]));
                                        a.grp_id[my_base + cnt] = tile * (TK_N / 8) + c * (HC / 8) + ch * 4 + gi;
                                    }
                                    cnt++;
                                }
                            }
                        }
                    };
                    tmem_ld_wait(va);
                    tmem_ld_wait(vb);
                    process(va, 0);
                    tmem_ld32(taddr + 64, va);
                    process(vb, 1);
                    tmem_ld32(taddr + 96, vb);
                    tmem_ld_wait(va);
                    tmem_ld_wait(vb);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[h]);  // the accumulator is in registers: release it early
                    process(va, 2);
                    process(vb, 3);
                    if (MODE == MODE_MAX && a.gpt != 8) a.maxes[((size_t)i * 2 + c) * a.ub + ub] = mx;
                }
                if (MODE == MODE_CAND) a.grp_cnt[(size_t)c * a.ub + ub] = cnt;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- prep: fp32 rows -> bf16 rows (stride kp, zero padded), norms, NaN flags ---------------------------
// rows_src == nullptr: row i of the output is row i of M; else row i is row rows_src[i] (gather of users).
//
// CENTRING (both sides).  With q_mean, p_mean the mean rows: d_i = fl(q_i - q_mean), e_u = fl(p_u - p_mean) and
//     <p_u, q_i> = <p_u, q_mean>  +  <p_mean, d_i>  +  <e_u, d_i>
//                  c_u: the same for all items of a user -- no comparison inside a user's list sees it
//                                  a_i: an item bias, computed in fp32 (bias_out), the same for all users
//                                                  what the bf16 GEMM computes, error <= 2^-7 |e_u||d_i|
// The factors of a trained model share a large common component (every score of a user sits near the mean rating, a user
// with few ratings is still close to the initial vector all users start from): against |p_u||q_i| the gaps between
// neighbouring scores cannot be resolved in bf16 -- measured on factors trained at the 1M x 625k shape: > 1 500 candidates
// per user and the exact per-user path for most users without centring.  The items are fed to the GEMM SORTED by a_i
// (rows_src = the permutation), so that the 32-item groups of the epilogue have a narrow bias range [amin, amax].
// What the bounds must cover: the exact score is s_ui = fl(<p_u, q_i>) = c_u + a_i + g_ui + R with g_ui the GEMM's value and
//     |R| <= 2^-7 * 1.05 |e_u||d_i|                                  (bf16 rounding of both operands, fp32 accumulation)
//          + gamma_k (|p_u||q_i| + |p_mean||d_i|)                    (sequential fp32 sum of the exact score; fp32 sum of a_i)
//          + 2^-24 (|p_u| + |e_u|)|d_i|                              (rounding of the two subtractions)
// gamma_k <= 2^-17 for k <= 128.  With the "norms" N_i = |d_i| + 2^-5 (|q_i| + |q_mean|) and eps_u = 1.05 * 2^-7 *
// (|e_u| + 2^-5 (|p_u| + |p_mean|)) the product eps_u N_i covers all of it (|d_i| <= |q_i| + |q_mean|, |e_u| <= |p_u| + |p_mean|).
// centre == nullptr: no centring (norm = |x|); other != nullptr: bias_out[i] = <other, x - centre>.
__global__ void __launch_bounds__(256)
k_topk_prep(const float *__restrict__ M, int m_rows, int k, const int *__restrict__ rows_src, int out_rows,
            int out_rows_padded, int kp, __nv_bfloat16 *out, float *norm, int *is_nan, const float *__restrict__ centre,
            const float *__restrict__ other, float *bias_out) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    float cn = 0.f;  // |centre|
    if (centre) {
        for (int d = lane; d < k; d += 32) cn += centre[d] * centre[d];
        for (int o = 16; o > 0; o >>= 1) cn += __shfl_xor_sync(kFullMask, cn, o);
        cn = sqrtf(cn);
    }
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < out_rows_padded; i += warps) {
        const int src = i < out_rows ? (rows_src ? rows_src[i] : i) : -1;
        const bool ok = src >= 0 && src < m_rows;
        float ss = 0.f, sc = 0.f, sb = 0.f;
        bool nan = false;
        for (int d = lane; d < kp; d += 32) {
            float x = 0.f;
            if (ok && d < k) x = M[(size_t)src * k + d];
            if (isnan(x)) nan = true;
            ss += x * x;
            if (centre && d < k) {
                const float c = x - centre[d];
                sc += c * c;
                if (other) sb = fmaf(other[d], c, sb);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(kFullMask, ss, o);
            sc += __shfl_xor_sync(kFullMask, sc, o);
            sb += __shfl_xor_sync(kFullMask, sb, o);
        }
        nan = __any_sync(kFullMask, nan) || !(ss <= 3.0e38f);  // overflowing rows are treated like NaN rows
        for (int d = lane; d < kp; d += 32) {
            float x = 0.f;
            if (ok && !nan && d < k) x = M[(size_t)src * k + d] - (centre ? centre[d] : 0.f);
            out[(size_t)i * kp + d] = __float2bfloat16_rn(x);
        }
        if (lane == 0) {
            const float nrm = centre ? sqrtf(sc) * 1.0000005f + (sqrtf(ss) + cn) * (1.0000005f / 32.f) : sqrtf(ss) * 1.0000005f;
            norm[i] = (ok && !nan) ? nrm : 0.f;  // rounded up a little
            is_nan[i] = (!ok || nan) ? 1 : 0;
            if (bias_out) bias_out[i] = (ok && !nan) ? sb : 0.f;
        }
    }
}

// sort key of an item: its bias a_i (rows with NaN / overflow: +inf, they go to the end); ids[i] = i; the NaN flags in
// ORIGINAL item order with their per-256 counts (for k_topk_nan_list: the lowest-numbered NaN items join every list)
__global__ void __launch_bounds__(256)
k_topk_item_keys(const float *__restrict__ Q, int n, int n_padded, int k, const float *__restrict__ centre,
                 const float *__restrict__ other, float *key, int *ids, int *is_nan_orig, int *blk_nan_orig) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // a warp per item; 8 items per block
    bool bad = true;
    if (i < n_padded) {
        float ss = 0.f, sb = 0.f;
        bool nan = false;
        if (i < n)
            for (int d = lane; d < k; d += 32) {
                const float x = Q[(size_t)i * k + d];
                nan |= isnan(x);
                ss += x * x;
                sb = fmaf(other[d], x - centre[d], sb);
            }
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(kFullMask, ss, o);
            sb += __shfl_xor_sync(kFullMask, sb, o);
        }
        bad = i >= n || __any_sync(kFullMask, nan) || !(ss <= 3.0e38f) || !(fabsf(sb) <= 3.0e38f);
        if (lane == 0) {
            key[i] = bad ? __int_as_float(0x7f800000) : sb;
            ids[i] = i;
            is_nan_orig[i] = bad ? 1 : 0;
        }
    }
    (void)blk_nan_orig;
}
__global__ void __launch_bounds__(256) k_topk_count_nan(const int *__restrict__ is_nan, int n, int *blk_nan) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = __syncthreads_count(i < n && is_nan[i] != 0);
    if (threadIdx.x == 0) blk_nan[blockIdx.x] = c;
}

// the centre of the item rows: mean over the rows without NaN / overflow (any vector would be correct; the mean makes the
// |d_i| small).  acc[kp] doubles + acc[kp] = number of rows counted; k_topk_centre_finish turns them into floats.
__global__ void __launch_bounds__(256)
k_topk_centre_sum(const float *__restrict__ M, int rows, int k, int kp, double *acc) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    double part[4] = {0.0, 0.0, 0.0, 0.0};  // kp <= 128: four dimensions per lane
    double cnt = 0.0, sq = 0.0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < rows; i += warps) {
        float x[4];
        float ss = 0.f;
        bool nan = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int d = lane + 32 * j;
            x[j] = d < k ? M[(size_t)i * k + d] : 0.f;
            nan |= isnan(x[j]);
            ss += x[j] * x[j];
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFullMask, ss, o);
        nan = __any_sync(kFullMask, nan) || !(ss <= 3.0e38f);
        if (!nan) {
#pragma unroll
            for (int j = 0; j < 4; j++) part[j] += (double)x[j];
            cnt += 1.0;
            sq += (double)ss;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (lane + 32 * j < kp && part[j] != 0.0) atomicAdd(acc + lane + 32 * j, part[j]);
    if (lane == 0 && cnt != 0.0) {
        atomicAdd(acc + kp, cnt);
        atomicAdd(acc + kp + 1, sq);  // sum of |row|^2: mean |row - mean|^2 = mean |row|^2 - |mean|^2
    }
}

// full centring: a second lower bound of the user's topk-th best (centred) score, from the item biases alone.  Positions
// [n_good - topk, n_good) of the bias-sorted order hold the topk items with the largest bias; each of them scores at least
// a_i - |e_u||d_i| - (GEMM-independent slack), so   tau_u >= a_(topk) - pnorm_u * (1 + eps) * max N_i   (pnorm_u >= |e_u|,
// N_i >= |d_i|, and eps_u N_i covers the fp32 terms of the bound in k_topk_prep).  Tight exactly where the chunk maxima
// of the sampled GEMM pass are not: for a user close to the mean user the best items are the items with the largest bias,
// which sit in a handful of neighbouring chunks.
__global__ void k_topk_top_bias(const float *__restrict__ a_cand, const float *__restrict__ qn_cand, const int *__restrict__ blk_nan,
                                int n_blocks, int n, int topk, float *out2) {
    __shared__ int s_bad;
    __shared__ float s_max[32];
    if (threadIdx.x == 0) {
        int bad = 0;
        for (int j = 0; j < n_blocks; j++) bad += blk_nan[j];
        s_bad = bad;
    }
    __syncthreads();
    const int n_good = n - s_bad;
    float dmax = 0.f;
    if (n_good >= topk)
        for (int i = n_good - topk + (int)threadIdx.x; i < n_good; i += blockDim.x) dmax = fmaxf(dmax, qn_cand[i]);
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(kFullMask, dmax, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = dmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) dmax = fmaxf(dmax, s_max[w]);
        out2[0] = n_good >= topk ? a_cand[n_good - topk] : __int_as_float(0xff800000);  // a_(topk); -inf: no such bound
        out2[1] = dmax;
    }
}
// full centring: the items of the last (highest-bias) tiles were dumped group by group with their GEMM values
// (k_topk_gemm<MODE_CAND> with a threshold of -inf); their per-item lower bounds g + a_i - eps_u N_i become `rows` more
// rows of `maxes` (one thread per user, [row][ub] like the chunk maxima).  The sampled pass leaves those tiles out, so no
// item is counted twice.
__global__ void __launch_bounds__(256)
k_topk_top_items(const float4 *__restrict__ grp_sc, const int *__restrict__ grp_id, const int *__restrict__ grp_cnt, int gmax,
                 int ub, const float *__restrict__ eps_arr, const float *__restrict__ a_cand, const float *__restrict__ qn_cand,
                 float *rows_out, int rows, int users_done, int n_items_padded) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= ub) return;
    const float eps = eps_arr[u], ninf = __int_as_float(0xff800000);
    int r = 0;
    for (int half = 0; half < 2 && u < users_done; half++) {  // (users past the batch's tiles have no lists)
        const int cnt = min(grp_cnt[(size_t)half * ub + u], gmax / 2);
        for (int gi = 0; gi < cnt && r + 8 <= rows; gi++) {
            const size_t at = (size_t)u * gmax + (size_t)half * (gmax / 2) + gi;
            const int item0 = grp_id[at] * 8;
            if (item0 < 0 || item0 + 8 > n_items_padded) continue;
            const float4 sa = grp_sc[at * 2], sb = grp_sc[at * 2 + 1];
            const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float av = a_cand[item0 + e], nv = qn_cand[item0 + e];
                float lb = fmaf(-eps, nv, sv[e] + av) - 4e-7f * (fabsf(sv[e]) + fabsf(av));
                if (!(lb == lb) || !(nv > -3.0e38f)) lb = ninf;  // a NaN / padding item gives no bound
                rows_out[(size_t)(r + e) * ub + u] = lb;
            }
            r += 8;
        }
    }
    for (; r < rows; r++) rows_out[(size_t)r * ub + u] = ninf;
}
__global__ void k_topk_fill(float *x, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void __launch_bounds__(256)
k_topk_tau_bias(float *tau, const float *__restrict__ pnorm, const int *__restrict__ p_nan, int ub, const float *__restrict__ top2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ub || p_nan[i]) return;
    const float a_top = top2[0], d_top = top2[1];
    if (!(a_top > -3.0e38f)) return;
    const float t = a_top - pnorm[i] * d_top * (1.0f + 1.01f * TK_EPS) - 1e-6f * fabsf(a_top);
    const float cur = tau[i];
    if (cur == cur && cur < 3.0e38f && t > cur) tau[i] = t;  // (+inf: a user without candidates stays that way)
}
__global__ void k_topk_centre_finish(const double *__restrict__ acc, int k, int kp, float *centre) {
    const int d = threadIdx.x;
    if (d < kp) {
        const double c = acc[kp];
        const float v = (d < k && c > 0.0) ? (float)(acc[d] / c) : 0.f;
        centre[d] = (v == v && fabsf(v) < 1.0e30f) ? v : 0.f;
    }
}

// item side: |q_v| per item with -inf for NaN / padding items (never candidates); per group of 32 items the largest
// norm for the two GEMM passes (+inf in the bound pass if the group holds a NaN or padding item: no lower bound from
// it); per block of 256 items the number of NaN items (for k_topk_nan_list)
__global__ void __launch_bounds__(256)
k_topk_item_bounds(const float *__restrict__ norm, const int *__restrict__ is_nan, int n, int n_padded, float *qn_cand,
                   float *qg_max, float *qg_cand, int *blk_nan, const float *__restrict__ bias, float *ag_min, float *ag_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // n_padded is a multiple of 256: whole warps, whole blocks
    const bool bad = is_nan[i] != 0;
    float vmax = bad ? __int_as_float(0x7f800000) : norm[i];
    float vcand = bad ? __int_as_float(0xff800000) : norm[i];
    // bias range of the 32-item group over its good items (a group with a bad item gives no lower bound anyway: its
    // qg_max is +inf; amin then only has to be finite)
    const float bv = bias ? bias[i] : 0.f;
    float amin = bad ? __int_as_float(0x7f800000) : bv, amax = bad ? __int_as_float(0xff800000) : bv;
    qn_cand[i] = vcand;
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(kFullMask, vmax, o));
        vcand = fmaxf(vcand, __shfl_xor_sync(kFullMask, vcand, o));
        amin = fminf(amin, __shfl_xor_sync(kFullMask, amin, o));
        amax = fmaxf(amax, __shfl_xor_sync(kFullMask, amax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        qg_max[i >> 5] = vmax;
        qg_cand[i >> 5] = vcand;
        ag_min[i >> 5] = amin <= 3.0e38f ? amin : 0.f;
        ag_max[i >> 5] = amax;  // -inf for a group without a good item: its threshold becomes +inf
    }
    const int c = __syncthreads_count(bad && i < n);
    if (threadIdx.x == 0) blk_nan[blockIdx.x] = c;
}
// the first `want` NaN items in ascending order: one block scans the per-256-item counts of k_topk_item_bounds
// (1024 at a time) and walks only the 256-item blocks that hold one
__global__ void __launch_bounds__(1024)
k_topk_nan_list(const int *__restrict__ is_nan, const int *__restrict__ blk_nan, int n_blocks, int n, int want, int *list,
                int *count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int bi = b0 + threadIdx.x;
        const int c = bi < n_blocks ? blk_nan[bi] : 0;
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        int slot = s_base + incl - c;
        for (int j = 0; j < w; j++) slot += s_warp[j];
        if (c > 0 && slot < want)
            for (int i = bi * 256; i < bi * 256 + 256 && i < n && slot < want; i++)
                if (is_nan[i]) list[slot++] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = s_base;
            for (int j = 0; j < 32; j++) t += s_warp[j];
            s_base = t;
        }
        __syncthreads();
        if (s_base >= want) break;
    }
    if (threadIdx.x == 0) *count = min(s_base, want);
}
__global__ void __launch_bounds__(256)
k_topk_user_eps(const float *__restrict__ norm, const int *__restrict__ is_nan, int ub, float *eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ub) eps[i] = is_nan[i] ? 0.f : norm[i] * TK_EPS;
}

// ---- tau_u = the topk-th largest of the user's tile maxima (a warp per user) --------------------------------
__device__ __forceinline__ unsigned ord_key(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
template <int MAXV>  // a warp per user; every lane keeps up to MAXV maxima in registers
__global__ void __launch_bounds__(256)
k_topk_tau(const float *__restrict__ maxes, int n_tiles, int ub, int users, const int *__restrict__ user_nan, int topk,
           float *tau) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= ub) return;
    float out = __int_as_float(0x7f800000);  // +inf: padding and NaN users collect nothing
    if (u < users && !user_nan[u]) {
        if (n_tiles < topk) {
            out = __int_as_float(0xff800000);  // not enough maxima for a bound: everything is a candidate
        } else {
            unsigned key[MAXV];
#pragma unroll
            for (int j = 0; j < MAXV; j++) {
                const int t = lane + 32 * j;
                key[j] = t < n_tiles ? ord_key(maxes[(size_t)t * ub + u]) : 0u;
            }
            unsigned prefix = 0;  // bitwise search of the largest key with at least topk maxima >= it
            for (int bit = 31; bit >= 0; bit--) {
                const unsigned cand = prefix | (1u << bit);
                int c = 0;
#pragma unroll
                for (int j = 0; j < MAXV; j++) c += key[j] >= cand ? 1 : 0;
                for (int t = lane + 32 * MAXV; t < n_tiles; t += 32)  // beyond the register budget: re-read
                    c += ord_key(maxes[(size_t)t * ub + u]) >= cand ? 1 : 0;
                c = __reduce_add_sync(kFullMask, c);
                if (c >= topk) prefix = cand;
            }
            const unsigned b = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
            out = __uint_as_float(b);
            if (isnan(out)) out = __int_as_float(0xff800000);
        }
    }
    if (lane == 0) tau[u] = out;
}

// ---- exact re-score + selection -------------------------------------------------------------------------------
// mf_predict (mf/mf.cpp:4295-4314): z = sum in index order from 0.0f, product rounded before the add; NaN -> b.
__device__ __forceinline__ float predict_exact_row(const float *__restrict__ p, const float *__restrict__ q, int k, float b) {
    float z = 0.0f;
    for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(p[d], q[d]));
    return isnan(z) ? b : z;
}
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {  // score desc, id asc
    return sa != sb ? sa > sb : ia < ib;
}

// The kth largest of vals[0..total) (total >= kth >= 1), for all threads of the block: radix select on the
// order-preserving keys, 8 bits a round, on a 256-bin shared-memory histogram; the scan of the bins (largest digit
// first) is done by warp 0.  Ends with a __syncthreads().
__device__ __forceinline__ float block_kth_largest(const float *vals, int total, int kth, int *s_hist, unsigned *s_prefix,
                                                   int *s_remaining) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        *s_prefix = 0u;
        *s_remaining = kth;
    }
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        const unsigned prefix = *s_prefix;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const unsigned key = ord_key(vals[i]);
            if (shift == 24 || (key >> (shift + 8)) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (warp == 0) {  // lane l owns digits 255-8l .. 248-8l
            const int remaining = *s_remaining;  // read by every lane before the shuffles, written after them
            int h[8], mine = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                h[t] = s_hist[255 - 8 * lane - t];
                mine += h[t];
            }
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= o) incl += t;
            }
            int before_me = incl - mine;  // keys with a larger digit than any of mine
            if (before_me < remaining && incl >= remaining) {  // the kth key of the bucket has one of my digits
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (before_me < remaining && before_me + h[t] >= remaining) {
                        *s_prefix = (prefix << 8) | (unsigned)(255 - 8 * lane - t);
                        *s_remaining = remaining - before_me;
                        before_me = remaining;  // done
                    } else {
                        before_me += h[t];
                    }
                }
            }
        }
        __syncthreads();
    }
    const unsigned kstar = *s_prefix;
    return __uint_as_float((kstar & 0x80000000u) ? (kstar & 0x7fffffffu) : ~kstar);
}

// Selection kernel, one block per user:
//   0. expand (GEMM path only): the GEMM pass left 8-item groups with their bf16-GEMM scores; an item is a candidate
//      if its upper bound s + eps_u|q_v| reaches tau_u (the per-item test; NaN and padding items have |q| = -inf);
//   1. prune: the exact score of a candidate lies in [s - eps_u|q_v|, s + eps_u|q_v|].  tau' = the topk-th largest
//      lower bound is again a lower bound of the exact topk-th score -- much tighter than the tile-maxima bound the
//      GEMM pass worked with -- and only candidates whose upper bound reaches it are kept;
//   2. exact scores of the survivors in the reference's summation order: the rows are staged through shared memory
//      in chunks of `rows` rows -- a warp per row, 16-byte cp.async when k is a multiple of 4 (one instruction per
//      512-byte row at k=128), row stride an odd number of 16-byte units so that the 128-bit column walk of 32
//      threads is conflict-free -- then one thread per candidate does the sequential fp32 sum
//      z = (...((0 + p0 q0) + p1 q1) + ...) of mf_predict;
//   3. the topk-th largest exact score, the (few more than topk) entries that reach it, a bitonic sort of those by
//      (score desc, id asc), first topk out.
// Blocks are small (128 threads, ~52 KB) so that four of them share an SM and one block's gather overlaps another's
// arithmetic.
constexpr int TK_SEL_THREADS = 128;
constexpr int TK_SEL_ROWS = 64;
template <int SZ, int VEC>  // SZ: capacity (power of two) of the candidate list and of the shared-memory sort;
                           // VEC: 8 = k % 8 == 0, rows read straight from global memory; 4 / 1 = staged rows
__global__ void __launch_bounds__(TK_SEL_THREADS)
k_topk_select(const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k, float b,
              const int *__restrict__ users, int nusers, int user0, const float4 *__restrict__ grp_sc,
              const int *__restrict__ grp_id, const int *__restrict__ grp_cnt, int gmax, int ub,
              const float *__restrict__ eps_arr, const float *__restrict__ tau_arr, const float *__restrict__ qn_cand,
              const float *__restrict__ a_cand, const int *__restrict__ perm, const int *__restrict__ nan_list, const int *__restrict__ nan_count, int all_items, int topk, int rows,
              int stride, int prune, int centred, int stage_floats, const int *__restrict__ sel_list, int *ovf_batch,
              int *idx_out, float *score_out, int *overflow, unsigned long long *stats) {
    // Two tiers share this kernel.  Tier 1 (SZ = 2048, sel_list == nullptr): a block per user of the batch; a user whose
    // expanded candidate list does not fit SZ is put on ovf_batch ([0] = count, then positions inside the batch) and left
    // alone.  Tier 2 (SZ = 8192, sel_list = tier 1's ovf_batch): block j takes the j-th user of the list -- the long tail of
    // a trained model's score distribution stays on the tensor-core path; only what does not fit there either (or whose
    // group list overflowed in the GEMM pass) goes on `overflow` for the exact per-user path.
    __shared__ int s_hist[256];
    __shared__ unsigned s_prefix;
    __shared__ int s_remaining, s_count, s_count2, s_count3;
    extern __shared__ float4 s_dyn4[];  // [k] user row, then [rows][stride] candidate rows / the lists of steps 0-1 and 3
    float *s_p = reinterpret_cast<float *>(s_dyn4), *s_q = s_p + ((k + 3) & ~3);
    float *s_sc = s_q + stage_floats;                       // [SZ] exact scores
    int *s_id = reinterpret_cast<int *>(s_sc + SZ);         // [SZ] their items
    int ul = blockIdx.x;  // user inside the batch
    if (sel_list) {
        if ((int)blockIdx.x >= sel_list[0]) return;
        ul = sel_list[1 + blockIdx.x];
    }
    if (ul >= nusers) return;
    const int u = users[user0 + ul];
    const bool u_ok = u >= 0 && u < m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    bool u_nan = false;
    for (int d = threadIdx.x; d < k; d += blockDim.x) {
        const float x = u_ok ? P[(size_t)u * k + d] : 0.f;
        s_p[d] = x;
        u_nan |= isnan(x);
    }
    if (threadIdx.x == 0) {
        s_count = 0;
        s_count2 = 0;
        s_count3 = 0;
    }
    u_nan = __syncthreads_or(u_nan) != 0;
    // a NaN user row makes every score b: the exact answer is items 0..topk-1, whatever the candidates were
    // (and so does a user id outside [0, m): mf_predict returns b, mf/mf.cpp:4297-4299)
    if (!all_items && (!u_ok || u_nan)) {
        for (int j = threadIdx.x; j < topk; j += blockDim.x) {
            idx_out[(size_t)(user0 + ul) * topk + j] = j < n ? j : -1;
            if (score_out) score_out[(size_t)(user0 + ul) * topk + j] = j < n ? b : 0.f;
        }
        return;
    }

    int total, found = 0;
    bool recorded = false;  // (thread 0) this user is already on the overflow list
    if (all_items) {
        total = n;
        for (int i = threadIdx.x; i < total; i += blockDim.x) s_id[i] = i;
    } else {
        // ---- 0. expand the groups into candidates (item, lower bound, upper bound) ----
        int *t_item = reinterpret_cast<int *>(s_q);
        float *t_lb = s_q + SZ, *t_ub = s_q + 2 * SZ;
        const float eps = eps_arr[ul], tau = tau_arr[ul];
        int g0 = grp_cnt[ul], g1 = grp_cnt[ub + ul];
        if (g0 > gmax / 2 || g1 > gmax / 2) {
            if (sel_list) return;  // (tier 1 has put this user on the global list already)
            if (threadIdx.x == 0) {  // overflow[0] = number of records, overflow[1 + i] = position of the user in `users`
                overflow[1 + atomicAdd(overflow, 1)] = user0 + ul;
                recorded = true;
            }
            if (ovf_batch) return;   // a second tier exists: no point in working on a truncated list
            g0 = min(g0, gmax / 2);
            g1 = min(g1, gmax / 2);
        }
        for (int gi = threadIdx.x; gi < g0 + g1; gi += blockDim.x) {
            const size_t at = (size_t)ul * gmax + (gi < g0 ? gi : gmax / 2 + (gi - g0));
            const int item0 = grp_id[at] * 8;
            const float4 sa = grp_sc[at * 2], sb = grp_sc[at * 2 + 1];
            const float4 na = *reinterpret_cast<const float4 *>(qn_cand + item0);
            const float4 nb = *reinterpret_cast<const float4 *>(qn_cand + item0 + 4);
            float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
            const float nv[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
            float slack[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (a_cand) {  // centred: GEMM value + item bias (positions are places in the bias-sorted order)
                const float4 aa = *reinterpret_cast<const float4 *>(a_cand + item0);
                const float4 ab = *reinterpret_cast<const float4 *>(a_cand + item0 + 4);
                const float av[8] = {aa.x, aa.y, aa.z, aa.w, ab.x, ab.y, ab.z, ab.w};
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    slack[e] = 4e-7f * (fabsf(sv[e]) + fabsf(av[e]));  // rounding of the sum
                    sv[e] += av[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float hi = fmaf(eps, nv[e], sv[e]) + slack[e];
                if (hi >= tau) {
                    const int slot = atomicAdd(&s_count, 1);
                    if (slot < SZ) {
                        t_item[slot] = perm ? perm[item0 + e] : item0 + e;
                        t_lb[slot] = fmaf(-eps, nv[e], sv[e]) - slack[e];
                        t_ub[slot] = hi;
                    }
                }
            }
        }
        __syncthreads();
        found = s_count;
        const int cn = min(*nan_count, topk);
        if (found + cn > SZ) {
            if (ovf_batch) {  // tier 1: hand the user to tier 2
                if (threadIdx.x == 0) ovf_batch[1 + atomicAdd(ovf_batch, 1)] = ul;
                return;
            }
            if (threadIdx.x == 0 && !recorded) overflow[1 + atomicAdd(overflow, 1)] = user0 + ul;
            found = min(found, SZ - cn);
        }
        for (int i = threadIdx.x; i < cn; i += blockDim.x) {  // a NaN item scores exactly b
            t_item[found + i] = nan_list[i];
            // (centred GEMM scores live on another scale than b: the NaN items then carry no bound at all -- they are
            // always kept and never raise the pruning threshold)
            t_lb[found + i] = centred ? __int_as_float(0xff800000) : b;
            t_ub[found + i] = centred ? __int_as_float(0x7f800000) : b;
        }
        total = found + cn;
        __syncthreads();
        if (!prune || total <= topk) {
            for (int i = threadIdx.x; i < total; i += blockDim.x) s_id[i] = t_item[i];
        } else {
            // ---- 1. prune ----
            const float tau2 = block_kth_largest(t_lb, total, topk, s_hist, &s_prefix, &s_remaining);
            for (int i = threadIdx.x; i < total; i += blockDim.x)
                if (t_ub[i] >= tau2) s_id[atomicAdd(&s_count2, 1)] = t_item[i];
            __syncthreads();
            total = s_count2;
        }
    }
    __syncthreads();  // the lists of steps 0-1 live where the rows are staged next
    if (stats && threadIdx.x == 0) {
        atomicAdd(stats, (unsigned long long)found);
        atomicAdd(stats + 1, (unsigned long long)total);
    }

    // ---- 2. exact scores ----
    if (VEC == 8) {
        // one thread per candidate walks its row with 256-bit loads (a whole 32-byte sector each, four in flight)
        const int k8 = k >> 3;
        const float4 *p4 = reinterpret_cast<const float4 *>(s_p);
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int id = s_id[i];
            const bool ok = (unsigned)id < (unsigned)n;
            const float *row = Q + (size_t)(ok ? id : 0) * k;
            float z = 0.0f;
            for (int j = 0; j < k8; j += 4) {
                float v[4][8];
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (j + t < k8)
                        asm volatile("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=f"(v[t][0]), "=f"(v[t][1]), "=f"(v[t][2]), "=f"(v[t][3]), "=f"(v[t][4]), "=f"(v[t][5]),
                                       "=f"(v[t][6]), "=f"(v[t][7])
                                     : "l"(row + 8 * (j + t)));
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (j + t < k8) {
                        const float4 pa = p4[2 * (j + t)], pb = p4[2 * (j + t) + 1];
                        z = __fadd_rn(z, __fmul_rn(pa.x, v[t][0]));
                        z = __fadd_rn(z, __fmul_rn(pa.y, v[t][1]));
                        z = __fadd_rn(z, __fmul_rn(pa.z, v[t][2]));
                        z = __fadd_rn(z, __fmul_rn(pa.w, v[t][3]));
                        z = __fadd_rn(z, __fmul_rn(pb.x, v[t][4]));
                        z = __fadd_rn(z, __fmul_rn(pb.y, v[t][5]));
                        z = __fadd_rn(z, __fmul_rn(pb.z, v[t][6]));
                        z = __fadd_rn(z, __fmul_rn(pb.w, v[t][7]));
                    }
            }
            // NaN -> b (mf/mf.cpp:4305-4306); a user or item outside the model scores b (4297-4299)
            s_sc[i] = (isnan(z) || !u_ok || !ok) ? b : z;
        }
        __syncthreads();
    } else {
    const bool VEC4 = VEC == 4;
    const uint32_t sq_addr = smem_u32(s_q);
    const int k4 = k >> 2;
    for (int base = 0; base < total; base += rows) {
        const int cnt = min(rows, total - base);
        // coalesced: a warp per row, all rows of the chunk in flight
        for (int r = warp; r < cnt; r += nwarps) {
            const int id = s_id[base + r];
            const float *src = Q + (size_t)((unsigned)id < (unsigned)n ? id : 0) * k;
            const uint32_t dst = sq_addr + (uint32_t)(r * stride) * 4u;
            if (VEC4) {
                for (int d4 = lane; d4 < k4; d4 += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * d4), "l"(src + 4 * d4) : "memory");
            } else {
                for (int d = lane; d < k; d += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * d), "l"(src + d) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
            const int id = s_id[base + r];
            float z = 0.0f;
            if (VEC4) {
                const float4 *q4 = reinterpret_cast<const float4 *>(s_q + r * stride);
                const float4 *p4 = reinterpret_cast<const float4 *>(s_p);
#pragma unroll 4
                for (int j = 0; j < k4; j++) {
                    const float4 qq = q4[j], pp = p4[j];
                    z = __fadd_rn(z, __fmul_rn(pp.x, qq.x));
                    z = __fadd_rn(z, __fmul_rn(pp.y, qq.y));
                    z = __fadd_rn(z, __fmul_rn(pp.z, qq.z));
                    z = __fadd_rn(z, __fmul_rn(pp.w, qq.w));
                }
            } else {
                const float *q = s_q + r * stride;
                for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(s_p[d], q[d]));
            }
            // NaN -> b (mf/mf.cpp:4305-4306); a user or item outside the model scores b (4297-4299)
            s_sc[base + r] = (isnan(z) || !u_ok || (unsigned)id >= (unsigned)n) ? b : z;
        }
        __syncthreads();
    }
    }

    // ---- 3. the first topk by (score desc, id asc) ----
    float *o_sc = s_q;
    int *o_id = reinterpret_cast<int *>(s_q) + SZ;
    int cnt3 = total;
    if (total > topk) {
        const float cut = block_kth_largest(s_sc, total, topk, s_hist, &s_prefix, &s_remaining);
        for (int i = threadIdx.x; i < total; i += blockDim.x)
            if (s_sc[i] >= cut) {
                const int slot = atomicAdd(&s_count3, 1);
                o_sc[slot] = s_sc[i];
                o_id[slot] = s_id[i];
            }
        __syncthreads();
        cnt3 = s_count3;
    } else {
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            o_sc[i] = s_sc[i];
            o_id[i] = s_id[i];
        }
    }
    int sz = 32;
    while (sz < cnt3) sz <<= 1;  // sort only as much as there is
    for (int i = cnt3 + threadIdx.x; i < sz; i += blockDim.x) {
        o_sc[i] = __int_as_float(0xff800000);
        o_id[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int size = 2; size <= sz; size <<= 1)
        for (int stride2 = size >> 1; stride2 > 0; stride2 >>= 1) {
            for (int i = threadIdx.x; i < sz / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride2 - 1)), hi = lo + stride2;
                const bool up = (lo & size) == 0;  // ascending block: "before" first
                const float a_s = o_sc[lo], b_s = o_sc[hi];
                const int a_i = o_id[lo], b_i = o_id[hi];
                const bool swap = up ? before(b_s, b_i, a_s, a_i) : before(a_s, a_i, b_s, b_i);
                if (swap) {
                    o_sc[lo] = b_s; o_sc[hi] = a_s;
                    o_id[lo] = b_i; o_id[hi] = a_i;
                }
            }
            __syncthreads();
        }
    for (int j = threadIdx.x; j < topk; j += blockDim.x) {
        const bool ok = j < cnt3 && o_id[j] != 0x7fffffff;
        idx_out[(size_t)(user0 + ul) * topk + j] = ok ? o_id[j] : -1;
        if (score_out) score_out[(size_t)(user0 + ul) * topk + j] = ok ? o_sc[j] : 0.f;
    }
}

// ---- host helpers -----------------------------------------------------------------------------------------------
int getenv_flag(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}
// 2-D bf16 row-major [rows][kp]; box = 64 elements (128 bytes) x box_rows, 128-byte swizzle
int make_tmap(CUtensorMap *tm, const void *base, int rows, int kp, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return 1;
    const cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TK_KATOM, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0
               : 1;
}

template <int KATOMS>
size_t gemm_smem_bytes() {
    return 1024 + (size_t)2 * KATOMS * TK_M * 128 + (size_t)TK_STAGES * KATOMS * TK_N * 128 + 16 * 8;
}

template <int KATOMS, int MODE>
int launch_gemm(const CUtensorMap &tmP, const CUtensorMap &tmQ, const TopkGemmArgs &a, int sm_count, cudaStream_t st) {
    const size_t smem = gemm_smem_bytes<KATOMS>();
    cudaError_t e = cudaFuncSetAttribute(k_topk_gemm<KATOMS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = a.n_user_tiles < sm_count ? a.n_user_tiles : sm_count;
    k_topk_gemm<KATOMS, MODE><<<grid, TK_THREADS, smem, st>>>(tmP, tmQ, a);
    return (int)cudaGetLastError();
}

// tier: 1 = a block per user, lists of 2048 candidates (ovf_batch: where users with longer lists go, or nullptr: straight to
// `overflow`); 2 = a block per entry of sel_list, lists of 8192 candidates
int launch_select(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers, int user0,
                  const float4 *grp_sc, const int *grp_id, const int *grp_cnt, int gmax, int ub, const float *eps,
                  const float *tau, const float *qn_cand, const float *a_cand, const int *perm, const int *nan_list,
                  const int *nan_count, int all_items, int topk, int prune, int centred, int tier, const int *sel_list, int *ovf_batch, int *idx_out, float *score_out,
                  int *overflow, unsigned long long *stats, cudaStream_t st, int grid_blocks = 0) {
    const int SZ = tier == 2 ? 8192 : 2048;
    if (grid_blocks <= 0) grid_blocks = nusers;
    const int vec = (k & 7) == 0 ? 8 : (k & 3) == 0 ? 4 : 1;
    const int stride = vec >= 4 ? 4 * ((k >> 2) | 1) : (k | 1);  // floats; an odd number of 16-byte (4-byte) units
    const int rows = TK_SEL_ROWS;
    size_t stage = vec == 8 ? 0 : (size_t)rows * stride;      // floats: row staging, or the three lists of steps 0-1
    if (stage < 3 * (size_t)SZ) stage = 3 * (size_t)SZ;
    const size_t smem = (size_t)(((k + 3) & ~3) + stage + 2 * (size_t)SZ) * 4;
    if (smem > 200 * 1024) return (int)cudaErrorNotSupported;
    auto kern = tier == 2 ? (vec == 8 ? k_topk_select<8192, 8> : vec == 4 ? k_topk_select<8192, 4> : k_topk_select<8192, 1>)
                          : (vec == 8 ? k_topk_select<2048, 8> : vec == 4 ? k_topk_select<2048, 4> : k_topk_select<2048, 1>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid_blocks, TK_SEL_THREADS, smem, st>>>(P, Q, m, n, k, b, users, nusers, user0, grp_sc, grp_id, grp_cnt, gmax, ub, eps, tau,
                                               qn_cand, a_cand, perm, nan_list, nan_count, all_items, topk, rows, stride, prune, centred,
                                               (int)stage, sel_list, ovf_batch, idx_out, score_out, overflow, stats);
    return (int)cudaGetLastError();
}

// ---- exact path for ONE user: every item scored like mf_predict (mf/mf.cpp:4295-4314: sequential fp32 sum from 0.0f,
// product rounded before the add, NaN -> b), then a full stable sort by falling score (equal scores: rising id).  Used
// for users whose candidate list overflowed in the GEMM path (many items tied at the cut: an all-zero user row, duplicate
// item rows) and for shapes the GEMM path does not take (k > 128 or topk > 128 with more than 2048 items).
__global__ void __launch_bounds__(256)
k_topk_exact_scores(const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k, float b,
                    const int *__restrict__ users, int pos, float *scores, int *ids) {
    extern __shared__ float s_row[];
    const int u = users[pos];
    const bool ok = u >= 0 && u < m;
    for (int d = threadIdx.x; d < k; d += blockDim.x) s_row[d] = ok ? P[(size_t)u * k + d] : 0.f;
    __syncthreads();
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const float *q = Q + (size_t)v * k;
        float z = 0.0f;
        for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(s_row[d], q[d]));
        if (!ok || z != z) z = b;
        scores[v] = z;
        ids[v] = v;
    }
}
__global__ void k_topk_exact_take(const float *__restrict__ scores, const int *__restrict__ ids, int n, int topk, int pos,
                                  int *idx_out, float *score_out) {
    for (int j = threadIdx.x; j < topk; j += blockDim.x) {
        idx_out[(size_t)pos * topk + j] = j < n ? ids[j] : -1;
        if (score_out) score_out[(size_t)pos * topk + j] = j < n ? scores[j] : 0.f;
    }
}

}  // namespace

// ================================================================================================================
extern "C" {

size_t mfk_topk_exact_work_bytes(int n) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, cub_bytes, (const float *)nullptr, (float *)nullptr, (const int *)nullptr,
                                              (int *)nullptr, n);
    return 4 * (((size_t)n * 4 + 255) & ~(size_t)255) + cub_bytes + 256;
}

// the exact top-k of the user at position `pos` of `users` (device array), written to row `pos` of the outputs
int mfk_topk_exact_user(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int pos, int topk,
                        int *idx_out, float *score_out, void *work, size_t work_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (work_bytes < mfk_topk_exact_work_bytes(n)) return (int)cudaErrorInvalidValue;
    const size_t seg = ((size_t)n * 4 + 255) & ~(size_t)255;
    uint8_t *w = (uint8_t *)work;
    float *sc = (float *)w, *sc2 = (float *)(w + seg);
    int *id = (int *)(w + 2 * seg), *id2 = (int *)(w + 3 * seg);
    size_t cub_bytes = work_bytes - 4 * seg;
    k_topk_exact_scores<<<std::min((n + 255) / 256, 148 * 8), 256, (size_t)k * 4, st>>>(P, Q, m, n, k, b, users, pos, sc, id);
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(w + 4 * seg, cub_bytes, sc, sc2, id, id2, n, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    k_topk_exact_take<<<1, 256, 0, st>>>(sc2, id2, n, topk, pos, idx_out, score_out);
    return (int)cudaGetLastError();
}

int mfk_topk_max_candidates(void) { return 1536; }  // 8-item groups per user the GEMM pass may keep (two half lists)

// Everything on `stream`; P [m][k], Q [n][k] fp32 on the device; users [nusers] on the device; outputs on the device.
// work: caller-provided device scratch of mfk_topk_work_bytes(...) bytes.  overflow_dev: int[1 + nusers], zeroed by the
// caller; [0] counts the users whose candidate list overflowed, [1 + i] are their positions in `users` (the result of
// such a user is not guaranteed: the caller re-runs them through mfk_topk_exact_user).
static int topk_gpt(int n_samp, int topk) { return n_samp < 4 * topk ? 8 : 2; }  // maxima per tile: per 32-column chunk or per half tile

constexpr int kTopTiles = 2;  // full centring: item tiles (256 items each) at the top of the bias order that are looked at item by item

static size_t topk_sort_bytes(int npad) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const float *)nullptr, (float *)nullptr, (const int *)nullptr, (int *)nullptr, npad);
    return b;
}

// batch_users: users per GEMM batch, a multiple of 256 (two 128-user accumulators per CTA)
size_t mfk_topk_work_bytes(int n, int k, int batch_users, int sample_stride) {
    const int kp = ((k + TK_KATOM - 1) / TK_KATOM) * TK_KATOM;
    const size_t npad = ((size_t)n + TK_N - 1) / TK_N * TK_N;
    const size_t ub = ((size_t)batch_users + 2 * TK_M - 1) / (2 * TK_M) * (2 * TK_M);
    const size_t n_tiles = npad / TK_N, n_samp = (n_tiles + sample_stride - 1) / sample_stride;
    size_t b = 0;
    b += npad * kp * 2 + 3 * npad * 4 + 2 * (npad / 32) * 4 + (npad / 256) * 4 + 7 * 256;  // Q bf16, norm, is_nan, qn_cand, qg_max, qg_cand, counts
    b += ub * kp * 2 + 4 * ub * 4 + 5 * 256;                            // P bf16, norm, is_nan, eps, tau
    b += (n_samp * 8 + (size_t)kTopTiles * TK_N) * ub * 4 + ub * (size_t)mfk_topk_max_candidates() * 36 + 2 * ub * 4 + 4 * 256;  // maxes, group lists, counts
    b += 1024 * 4 + 64 + 64 + 3 * 256;                                  // nan list, counters, stats
    b += 2 * 130 * 8 + 2 * 128 * 4 + 64 + (ub + 1) * 4 + 5 * 256;            // centres of both sides, tier-2 list
    b += 6 * npad * 4 + 2 * (npad / 32) * 4 + topk_sort_bytes((int)npad) + 10 * 256;  // bias keys, permutation, flags, sort scratch
    return b + 4096;
}

int mfk_topk(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers, int topk,
             int *idx_out, float *score_out, void *work, size_t work_bytes, int batch_users, int sample_stride,
             int sm_count, int *overflow_dev, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nusers <= 0) return 0;
    const int gmax = mfk_topk_max_candidates();
    // small item sets: every item is a candidate, no GEMM
    if (n + 0 <= 2048 && topk <= 2048) {
        return launch_select(P, Q, m, n, k, b, users, nusers, 0, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, 1, topk, 0, 0, 1, nullptr, nullptr, idx_out, score_out, overflow_dev, nullptr, st);
    }
    const int kp = ((k + TK_KATOM - 1) / TK_KATOM) * TK_KATOM;
    if (kp > 128 || topk > 128 || topk < 1) return (int)cudaErrorNotSupported;
    if (work_bytes < mfk_topk_work_bytes(n, k, batch_users, sample_stride)) return (int)cudaErrorInvalidValue;
    const int npad = (n + TK_N - 1) / TK_N * TK_N, ub = (batch_users + 2 * TK_M - 1) / (2 * TK_M) * (2 * TK_M);
    const int n_tiles = npad / TK_N, n_samp = (n_tiles + sample_stride - 1) / sample_stride;

    // carve the scratch
    uint8_t *w = (uint8_t *)work;
    auto take = [&](size_t bytes) {
        uint8_t *p = w;
        w += (bytes + 255) & ~(size_t)255;
        return p;
    };
    __nv_bfloat16 *Qb = (__nv_bfloat16 *)take((size_t)npad * kp * 2);
    float *qnorm = (float *)take((size_t)npad * 4);
    int *q_nan = (int *)take((size_t)npad * 4);
    float *qn_cand = (float *)take((size_t)npad * 4);
    int *blk_nan = (int *)take((size_t)(npad / 256) * 4);
    float *qg_max = (float *)take((size_t)(npad / 32) * 4), *qg_cand = (float *)take((size_t)(npad / 32) * 4);
    __nv_bfloat16 *Pb = (__nv_bfloat16 *)take((size_t)ub * kp * 2);
    float *pnorm = (float *)take((size_t)ub * 4);
    int *p_nan = (int *)take((size_t)ub * 4);
    float *eps = (float *)take((size_t)ub * 4), *tau = (float *)take((size_t)ub * 4);
    const int gpt = topk_gpt(n_samp, topk);
    float *maxes = (float *)take(((size_t)n_samp * gpt + (size_t)kTopTiles * TK_N) * ub * 4);
    float4 *grp_sc = (float4 *)take((size_t)ub * gmax * 32);
    int *grp_id = (int *)take((size_t)ub * gmax * 4);
    int *grp_cnt = (int *)take((size_t)2 * ub * 4);
    int *nan_list = (int *)take(1024 * 4), *nan_count = (int *)take(64);
    unsigned long long *stats = (unsigned long long *)take(64);
    double *centre_acc = (double *)take(sizeof(double) * 2 * 130);
    float *top2 = (float *)take(64);
    float *centre = (float *)take(sizeof(float) * 128), *centre_p = (float *)take(sizeof(float) * 128);
    float *a_key = (float *)take((size_t)npad * 4), *a_key_sorted = (float *)take((size_t)npad * 4), *a_cand = (float *)take((size_t)npad * 4);
    int *ids = (int *)take((size_t)npad * 4), *perm = (int *)take((size_t)npad * 4), *nan_orig = (int *)take((size_t)npad * 4);
    float *ag_min = (float *)take((size_t)(npad / 32) * 4), *ag_max = (float *)take((size_t)(npad / 32) * 4);
    const size_t sort_bytes = topk_sort_bytes((int)npad);
    void *sort_tmp = take(sort_bytes);
    int *ovf_batch = (int *)take(sizeof(int) * ((size_t)ub + 1));
    static const bool want_stats = getenv_flag("MFB200_TOPK_STATS", 0) != 0;
    static const int prune = getenv_flag("MFB200_TOPK_PRUNE", 1);
    if (want_stats) cudaMemsetAsync(stats, 0, 64, st);

    // Centring (see k_topk_prep).  Level 0: none.  1: the items are centred on their mean row.  2: both sides are centred,
    // the item bias travels beside the GEMM and the items are fed to it sorted by bias.  Level 2 pays when the users are
    // close to their mean row -- a trained model; it costs when they are not (the best items of a user then spread over
    // fewer chunks of the sampled pass, whose maxima give a lower threshold).  MFB200_TOPK_CENTRE=0/1/2 forces a level;
    // default: 2 when |mean user| >= 2 * rms |user - mean user|, else 1.
    const int centre_env = getenv_flag("MFB200_TOPK_CENTRE", -1);  // (read per call: a host may change it between calls)
    int centred = centre_env < 0 ? 1 : (centre_env > 2 ? 2 : centre_env);
    if (centred) {
        cudaMemsetAsync(centre_acc, 0, sizeof(double) * 2 * 130, st);
        k_topk_centre_sum<<<148 * 4, 256, 0, st>>>(Q, n, k, kp, centre_acc);
        k_topk_centre_finish<<<1, 128, 0, st>>>(centre_acc, k, kp, centre);
        if (centre_env < 0 || centred == 2) {
            k_topk_centre_sum<<<148 * 4, 256, 0, st>>>(P, m, k, kp, centre_acc + 130);
            k_topk_centre_finish<<<1, 128, 0, st>>>(centre_acc + 130, k, kp, centre_p);
        }
        if (centre_env < 0) {
            double hacc[130];
            if (cudaMemcpyAsync(hacc, centre_acc + 130, sizeof(double) * (size_t)(kp + 2), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess)
                return (int)cudaGetLastError();
            const double cnt = hacc[kp];
            double mean2 = 0.0;
            for (int d = 0; d < k; d++) mean2 += (hacc[d] / (cnt > 0 ? cnt : 1.0)) * (hacc[d] / (cnt > 0 ? cnt : 1.0));
            const double var = cnt > 0 ? hacc[kp + 1] / cnt - mean2 : 0.0;  // mean |p_u - p_mean|^2
            centred = (cnt > 0 && mean2 >= 4.0 * (var > 0 ? var : 0.0)) ? 2 : 1;
        }
    }
    // pass A on every third tile when the items are many and not sorted by bias (measured at 500k items, top-100, random
    // factors: 5.5M -> 6.0M users/s, 433 -> 633 candidates per user; with sorted items the coarser sample costs more than it
    // saves); MFB200_TOPK_STRIDE keeps the caller's choice
    if (centred != 2 && sample_stride == 2 && n_tiles >= 18 * topk && !getenv("MFB200_TOPK_STRIDE")) sample_stride = 3;
    if (centred == 2) {
        k_topk_item_keys<<<npad / 8, 256, 0, st>>>(Q, n, npad, k, centre, centre_p, a_key, ids, nan_orig, nullptr);
        size_t sb = sort_bytes;
        cudaError_t se = cub::DeviceRadixSort::SortPairs(sort_tmp, sb, a_key, a_key_sorted, ids, perm, npad, 0, 32, st);
        if (se != cudaSuccess) return (int)se;
        k_topk_prep<<<148 * 8, 256, 0, st>>>(Q, n, k, perm, npad, npad, kp, Qb, qnorm, q_nan, centre, centre_p, a_cand);
        k_topk_item_bounds<<<npad / 256, 256, 0, st>>>(qnorm, q_nan, npad, npad, qn_cand, qg_max, qg_cand, blk_nan, a_cand, ag_min, ag_max);
        k_topk_count_nan<<<npad / 256, 256, 0, st>>>(nan_orig, n, blk_nan);  // (original item order: the NaN list wants ids)
        k_topk_top_bias<<<1, 256, 0, st>>>(a_cand, qn_cand, blk_nan, npad / 256, n, topk, top2);
    } else {
        cudaMemsetAsync(ag_min, 0, sizeof(float) * (size_t)(npad / 32), st);
        cudaMemsetAsync(ag_max, 0, sizeof(float) * (size_t)(npad / 32), st);
        k_topk_prep<<<148 * 8, 256, 0, st>>>(Q, n, k, nullptr, n, npad, kp, Qb, qnorm, q_nan, centred ? centre : nullptr, nullptr, nullptr);
        k_topk_item_bounds<<<npad / 256, 256, 0, st>>>(qnorm, q_nan, n, npad, qn_cand, qg_max, qg_cand, blk_nan, nullptr, ag_min, ag_max);
    }
    k_topk_nan_list<<<1, 1024, 0, st>>>(centred == 2 ? nan_orig : q_nan, blk_nan, npad / 256, n, topk, nan_list, nan_count);
    CUtensorMap tmQ, tmP;
    if (make_tmap(&tmQ, Qb, npad, kp, TK_N) || make_tmap(&tmP, Pb, ub, kp, TK_M)) return (int)cudaErrorUnknown;

    for (int u0 = 0; u0 < nusers; u0 += batch_users) {
        const int nu = nusers - u0 < batch_users ? nusers - u0 : batch_users;
        // one accumulator per CTA while that still gives every SM a tile; two (the Q stream is then read once
        // per 256 users, half the L2 traffic per score) when there are more users than that
        const int nh = (nu + TK_M - 1) / TK_M <= sm_count ? 1 : 2;
        const int nu_pad = (nu + nh * TK_M - 1) / (nh * TK_M) * (nh * TK_M);
        k_topk_prep<<<148 * 4, 256, 0, st>>>(P, m, k, users + u0, nu, ub, kp, Pb, pnorm, p_nan, centred == 2 ? centre_p : nullptr, nullptr, nullptr);
        k_topk_user_eps<<<(ub + 255) / 256, 256, 0, st>>>(pnorm, p_nan, ub, eps);
        TopkGemmArgs a;
        a.n_user_tiles = nu_pad / (nh * TK_M);
        a.nh = nh;
        a.n_item_tiles = n_tiles;
        a.ub = ub;
        a.eps = eps;
        a.tau = tau;
        a.grp_sc = grp_sc;
        a.grp_id = grp_id;
        a.grp_cnt = grp_cnt;
        a.gmax = gmax;
        a.maxes = maxes;
        int rc = 0;
        // full centring: the last kTopTiles item tiles hold the items with the largest bias -- for a user near the mean user
        // that is where the whole answer sits, a handful of chunks whose maxima would say little.  Those tiles are dumped
        // item by item (candidate mode with a threshold of -inf) and left out of the sampled pass.
        const int top_tiles = centred == 2 ? (n_tiles > 2 * kTopTiles ? kTopTiles : 0) : 0;
        const int top_rows = top_tiles * TK_N;
        const int n_samp_b = ((n_tiles - top_tiles) + sample_stride - 1) / sample_stride;
        a.tile0 = 0;
        if (top_tiles) {
            k_topk_fill<<<(ub + 255) / 256, 256, 0, st>>>(tau, ub, -INFINITY);
            a.n_item_tiles = n_tiles;
            a.tile0 = n_tiles - top_tiles;
            a.tile_stride = 1;
            a.qg = qg_cand;
            a.ag = ag_max;
            rc = kp == 64 ? launch_gemm<1, MODE_CAND>(tmP, tmQ, a, sm_count, st) : launch_gemm<2, MODE_CAND>(tmP, tmQ, a, sm_count, st);
            if (rc) return rc;
            k_topk_top_items<<<(ub + 255) / 256, 256, 0, st>>>(grp_sc, grp_id, grp_cnt, gmax, ub, eps, a_cand, qn_cand,
                                                              maxes + (size_t)n_samp_b * gpt * ub, top_rows, nu_pad, npad);
            a.tile0 = 0;
        }
        // pass A: tile maxima of the lower bound on every sample_stride-th tile
        a.n_item_tiles = n_tiles - top_tiles;
        a.tile_stride = sample_stride;
        a.gpt = gpt;
        a.qg = qg_max;
        a.ag = ag_min;
        rc = kp == 64 ? launch_gemm<1, MODE_MAX>(tmP, tmQ, a, sm_count, st) : launch_gemm<2, MODE_MAX>(tmP, tmQ, a, sm_count, st);
        if (rc) return rc;
        a.n_item_tiles = n_tiles;
        const int n_rows = n_samp_b * gpt + top_rows;
        // (values beyond the register budget of an instantiation are re-read 32 times: 4.2 ms instead of 0.4 per batch when
        // 2 952 rows met the 2 048 of <64>)
        if (n_rows <= 32 * 16)
            k_topk_tau<16><<<(ub * 32 + 255) / 256, 256, 0, st>>>(maxes, n_rows, ub, nu, p_nan, topk, tau);
        else if (n_rows <= 32 * 64)
            k_topk_tau<64><<<(ub * 32 + 255) / 256, 256, 0, st>>>(maxes, n_rows, ub, nu, p_nan, topk, tau);
        else
            k_topk_tau<128><<<(ub * 32 + 255) / 256, 256, 0, st>>>(maxes, n_rows, ub, nu, p_nan, topk, tau);
        if (centred == 2) k_topk_tau_bias<<<(ub + 255) / 256, 256, 0, st>>>(tau, pnorm, p_nan, ub, top2);
        // pass C: candidates
        a.tile_stride = 1;
        a.qg = qg_cand;
        a.ag = ag_max;
        rc = kp == 64 ? launch_gemm<1, MODE_CAND>(tmP, tmQ, a, sm_count, st) : launch_gemm<2, MODE_CAND>(tmP, tmQ, a, sm_count, st);
        if (rc) return rc;
        cudaMemsetAsync(ovf_batch, 0, sizeof(int), st);
        rc = launch_select(P, Q, m, n, k, b, users, nu, u0, grp_sc, grp_id, grp_cnt, gmax, ub, eps, tau, qn_cand, centred == 2 ? a_cand : nullptr,
                           centred == 2 ? perm : nullptr, nan_list, nan_count, 0, topk, prune, centred, 1, nullptr, ovf_batch, idx_out, score_out, overflow_dev,
                           want_stats ? stats : nullptr, st);
        if (rc) return rc;
        // tier 2 for the users whose list did not fit 2048 entries: a block per entry of tier 1's list (its length is read
        // back: one stream synchronisation per batch of 37 888 users)
        int n_tier2 = 0;
        if (cudaMemcpyAsync(&n_tier2, ovf_batch, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return (int)cudaGetLastError();
        if (n_tier2 > 0) {
            rc = launch_select(P, Q, m, n, k, b, users, nu, u0, grp_sc, grp_id, grp_cnt, gmax, ub, eps, tau, qn_cand, centred == 2 ? a_cand : nullptr,
                               centred == 2 ? perm : nullptr, nan_list, nan_count, 0, topk, prune, centred, 2, ovf_batch, nullptr, idx_out, score_out, overflow_dev,
                               want_stats ? stats : nullptr, st, n_tier2);
            if (rc) return rc;
            if (want_stats) fprintf(stderr, "mfb200 topk stats: %d users of this batch went through the second tier\n", n_tier2);
        }
    }
    if (want_stats) {
        unsigned long long h[2] = {0, 0};
        cudaStreamSynchronize(st);
        cudaMemcpy(h, stats, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "mfb200 topk stats: %.1f candidate items per user after the GEMM pass, %.1f re-scored exactly\n",
                (double)h[0] / nusers, (double)h[1] / nusers);
    }
    return 0;
}

}  // extern "C"
