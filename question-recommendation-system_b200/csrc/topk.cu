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
#include <stdint.h>

#include "kernels.h"

namespace {

constexpr int TK_M = 128;        // users per tile  (UMMA M)
constexpr int TK_N = 256;        // items per tile  (UMMA N)
constexpr int TK_KATOM = 64;     // bf16 elements per 128-byte swizzle atom
constexpr int TK_STAGES = 2;     // Q tile stages in shared memory
constexpr int TK_THREADS = 320;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (2 per TMEM lane quarter)
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
    int n_user_tiles;   // tiles of 128 users in the batch
    int n_item_tiles;   // tiles of 256 items
    int tile_stride;    // MODE_MAX: every tile_stride-th item tile is sampled; MODE_CAND: 1
    int gpt;            // MODE_MAX: maxima per tile: 2 (per half tile = epilogue warp) or 8 (32-column chunks)
    int ub;             // users in the batch, padded to 128
    const float *eps;   // [ub] eps_u (0 for padding / NaN users)
    const float *qn;    // [n_item_tiles*256] |q_v|: +inf (MODE_MAX) / -inf (MODE_CAND) for NaN and padding items
    float *maxes;       // MODE_MAX: [n_sampled_tiles * gpt][ub]
    const float *tau;   // MODE_CAND: [ub] (+inf: no candidates)
    int *cand;          // MODE_CAND: [ub][cmax]
    int *cand_cnt;      // MODE_CAND: [2][ub]: per half list (cmax/2 entries each); may exceed: overflow
    int cmax;
};

template <int KATOMS, int MODE>
__global__ void __launch_bounds__(TK_THREADS, 1)
k_topk_gemm(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, const TopkGemmArgs a) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024-byte aligned
    constexpr uint32_t A_BYTES = KATOMS * TK_M * 128;          // 16 KB per atom
    constexpr uint32_t B_BYTES = KATOMS * TK_N * 128;          // 32 KB per atom
    uint8_t *sA = smem;
    uint8_t *sB = smem + A_BYTES;
    float *s_qn = reinterpret_cast<float *>(sB + TK_STAGES * B_BYTES);  // [8 warps][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_qn + 4 * TK_N);
    uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6, *a_full = bars + 8, *a_free = bars + 9;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 8);  // one arrival per epilogue warp
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

    const int n_it = (a.n_item_tiles + a.tile_stride - 1) / a.tile_stride;  // item tiles visited per user tile

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it_glob = 0, ut_count = 0;
            for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x, ut_count++) {
                if (ut_count > 0) mbar_wait(a_free, (ut_count - 1) & 1);  // the MMAs of the previous user tile are done
                mbar_expect_tx(a_full, A_BYTES);
                for (int ka = 0; ka < KATOMS; ka++) tma_load_2d(&tmP, sA + ka * (TK_M * 128), a_full, ka * TK_KATOM, ut * TK_M);
                for (int i = 0; i < n_it; i++, it_glob++) {
                    const int s = it_glob & 1;
                    mbar_wait(&empty[s], ((it_glob >> 1) & 1) ^ 1);
                    mbar_expect_tx(&full[s], B_BYTES);
                    const int tile = i * a.tile_stride;
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
                    const int s = it_glob & 1, buf = it_glob & 1;
                    mbar_wait(&full[s], (it_glob >> 1) & 1);
                    mbar_wait(&tempty[buf], ((it_glob >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)buf * TK_N;
#pragma unroll
                    for (int ka = 0; ka < KATOMS; ka++) {
                        const uint64_t ad = smem_desc_sw128(smem_u32(sA + ka * (TK_M * 128)));
                        const uint64_t bd = smem_desc_sw128(smem_u32(sB + s * B_BYTES + ka * (TK_N * 128)));
#pragma unroll
                        for (int k4 = 0; k4 < TK_KATOM / 16; k4++)  // UMMA K = 16 bf16 = 32 bytes inside the atom
                            umma_bf16(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, (ka | k4) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);    // the Q stage may be overwritten when these MMAs are done
                    umma_commit(&tfull[buf]);  // ... and the accumulator may be read
                }
                umma_commit(a_free);
            }
        }
    } else {
        // ===== epilogue: thread <-> user row; TMEM lane quarter = warp % 4; two warps share a quarter, each
        // takes one half (128 columns) of the tile =====
        const int q4 = warp & 3, half = (warp - 2) >> 2;
        const int row = q4 * 32 + lane;
        constexpr int HC = TK_N / 2;  // columns per warp
        float *my_qn = s_qn + (warp - 2) * HC;
        uint32_t it_glob = 0;
        for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x) {
            const int ub = ut * TK_M + row;
            const float eps = a.eps[ub];
            float tau = 0.f;
            int cnt = 0;  // MODE_CAND: candidates found by this thread in its half; lists are merged by slots below
            if (MODE == MODE_CAND) tau = a.tau[ub];
            // the two halves append to disjoint halves of the user's list (cmax/2 each); counts are kept per half
            int *my_cand = MODE == MODE_CAND ? a.cand + (size_t)ub * a.cmax + (size_t)half * (a.cmax / 2) : nullptr;
            const int my_cmax = a.cmax / 2;
            for (int i = 0; i < n_it; i++, it_glob++) {
                const int buf = it_glob & 1;
                const int tile = i * a.tile_stride;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < HC / 32; j++)
                    my_qn[lane + 32 * j] = __ldg(a.qn + (size_t)tile * TK_N + half * HC + lane + 32 * j);
                __syncwarp();
                mbar_wait(&tfull[buf], (it_glob >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)buf * TK_N + (uint32_t)half * HC;
                float mx = __int_as_float(0xff800000);  // -inf
                uint32_t va[32], vb[32];
                tmem_ld32(taddr, va);
                tmem_ld_wait(va);
                auto process = [&](const uint32_t *vr, int ch) {
                    float v[32];
#pragma unroll
                    for (int t = 0; t < 32; t++) v[t] = __uint_as_float(vr[t]);
                    const float4 *qv = reinterpret_cast<const float4 *>(my_qn + ch * 32);
                    if (MODE == MODE_MAX) {
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float4 qq = qv[j];
                            mx = fmaxf(mx, fmaf(-eps, qq.x, v[4 * j + 0]));
                            mx = fmaxf(mx, fmaf(-eps, qq.y, v[4 * j + 1]));
                            mx = fmaxf(mx, fmaf(-eps, qq.z, v[4 * j + 2]));
                            mx = fmaxf(mx, fmaf(-eps, qq.w, v[4 * j + 3]));
                        }
                        if (a.gpt == 8) {
                            a.maxes[((size_t)i * 8 + half * 4 + ch) * a.ub + ub] = mx;
                            mx = __int_as_float(0xff800000);
                        }
                    } else {
                        unsigned hit = 0u;  // bit c: column c of the chunk reaches tau (branch-free in the common case)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float4 qq = qv[j];
                            hit |= (fmaf(eps, qq.x, v[4 * j + 0]) >= tau ? 1u : 0u) << (4 * j + 0);
                            hit |= (fmaf(eps, qq.y, v[4 * j + 1]) >= tau ? 1u : 0u) << (4 * j + 1);
                            hit |= (fmaf(eps, qq.z, v[4 * j + 2]) >= tau ? 1u : 0u) << (4 * j + 2);
                            hit |= (fmaf(eps, qq.w, v[4 * j + 3]) >= tau ? 1u : 0u) << (4 * j + 3);
                        }
                        const int item0 = tile * TK_N + half * HC + ch * 32;
                        while (hit) {  // rare, and compact: no score is needed, only the column
                            const int c = __ffs(hit) - 1;
                            hit &= hit - 1u;
                            if (cnt < my_cmax) my_cand[cnt] = item0 + c;
                            cnt++;
                        }
                    }
                };
                // chunk ch+1 is loaded from TMEM while chunk ch is processed
                tmem_ld32(taddr + 32, vb);
                process(va, 0);
                tmem_ld_wait(vb);
                tmem_ld32(taddr + 64, va);
                process(vb, 1);
                tmem_ld_wait(va);
                tmem_ld32(taddr + 96, vb);
                process(va, 2);
                tmem_ld_wait(vb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);  // the accumulator is in registers: release it early
                process(vb, 3);
                if (MODE == MODE_MAX && a.gpt != 8) a.maxes[((size_t)i * 2 + half) * a.ub + ub] = mx;
            }
            if (MODE == MODE_CAND) a.cand_cnt[(size_t)half * a.ub + ub] = cnt;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- prep: fp32 rows -> bf16 rows (stride kp, zero padded), norms, NaN flags ---------------------------
// rows_src == nullptr: row i of the output is row i of M; else row i is row rows_src[i] (gather of users).
__global__ void __launch_bounds__(256)
k_topk_prep(const float *__restrict__ M, int m_rows, int k, const int *__restrict__ rows_src, int out_rows,
            int out_rows_padded, int kp, __nv_bfloat16 *out, float *norm, int *is_nan) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < out_rows_padded; i += warps) {
        const int src = i < out_rows ? (rows_src ? rows_src[i] : i) : -1;
        const bool ok = src >= 0 && src < m_rows;
        float ss = 0.f;
        bool nan = false;
        for (int d = lane; d < kp; d += 32) {
            float x = 0.f;
            if (ok && d < k) x = M[(size_t)src * k + d];
            if (isnan(x)) nan = true;
            ss += x * x;
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFullMask, ss, o);
        nan = __any_sync(kFullMask, nan) || !(ss <= 3.0e38f);  // overflowing rows are treated like NaN rows
        for (int d = lane; d < kp; d += 32) {
            float x = 0.f;
            if (ok && !nan && d < k) x = M[(size_t)src * k + d];
            out[(size_t)i * kp + d] = __float2bfloat16_rn(x);
        }
        if (lane == 0) {
            norm[i] = (ok && !nan) ? sqrtf(ss) * 1.0000005f : 0.f;  // rounded up a little
            is_nan[i] = (!ok || nan) ? 1 : 0;
        }
    }
}

// item side: the two norm arrays of the GEMM passes and the list of NaN items (ascending)
__global__ void __launch_bounds__(256)
k_topk_item_bounds(const float *__restrict__ norm, const int *__restrict__ is_nan, int n_padded, float *qn_max,
                   float *qn_cand) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    const bool bad = is_nan[i] != 0;
    qn_max[i] = bad ? __int_as_float(0x7f800000) : norm[i];   // lower bound -inf: never a maximum
    qn_cand[i] = bad ? __int_as_float(0xff800000) : norm[i];  // upper bound -inf: never a candidate
}
// the first `want` NaN items in ascending order: one block walks the flags 1024 at a time (ballot + prefix count)
__global__ void __launch_bounds__(1024)
k_topk_nan_list(const int *__restrict__ is_nan, int n, int want, int *list, int *count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool f = i < n && is_nan[i] != 0;
        const unsigned bal = __ballot_sync(kFullMask, f);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int before_me = s_base;
        for (int j = 0; j < w; j++) before_me += s_warp[j];
        const int slot = before_me + __popc(bal & ((1u << lane) - 1u));
        if (f && slot < want) list[slot] = i;
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
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFullMask, c, o);
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

// Exact scores of `total` candidates of one user, in the reference's summation order.  The candidate rows are
// staged through shared memory with coalesced loads (a warp per row, 128 bytes per instruction) in chunks of
// `rows` rows with an odd stride (conflict-free column walks); then one thread per candidate does the sequential
// fp32 sum  z = (...((0 + p0 q0) + p1 q1) + ...)  of mf_predict.
template <int SZ>  // SZ: capacity (power of two) of the shared-memory sort
__global__ void __launch_bounds__(256)
k_topk_select(const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k, float b,
              const int *__restrict__ users, int nusers, int user0, const int *__restrict__ cand,
              const int *__restrict__ cand_cnt, int cmax, int ub, const int *__restrict__ nan_list,
              const int *__restrict__ nan_count, int all_items, int topk, int rows, int *idx_out, float *score_out,
              int *overflow) {
    __shared__ float s_sc[SZ];
    __shared__ int s_id[SZ];
    extern __shared__ float s_dyn[];  // [k] user row, then [rows][stride] candidate rows
    const int stride = k | 1;
    float *s_p = s_dyn, *s_q = s_dyn + ((k + 3) & ~3);
    const int ul = blockIdx.x;  // user inside the batch
    if (ul >= nusers) return;
    const int u = users[user0 + ul];
    const bool u_ok = u >= 0 && u < m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // candidates: all items, or the two half lists of the GEMM pass followed by the lowest NaN items
    int c0 = 0, c1 = 0, cn = 0, total;
    if (all_items) {
        total = n;
    } else {
        c0 = cand_cnt[ul];
        c1 = cand_cnt[ub + ul];
        if (c0 > cmax / 2 || c1 > cmax / 2) {
            if (threadIdx.x == 0) atomicExch(overflow, 1);
            c0 = min(c0, cmax / 2);
            c1 = min(c1, cmax / 2);
        }
        cn = min(*nan_count, topk);
        total = c0 + c1 + cn;
    }
    int sz = 64;
    while (sz < total) sz <<= 1;  // sort only as much as there is
    bool u_nan = false;
    for (int d = threadIdx.x; d < k; d += blockDim.x) {
        const float x = u_ok ? P[(size_t)u * k + d] : 0.f;
        s_p[d] = x;
        u_nan |= isnan(x);
    }
    u_nan = __syncthreads_or(u_nan) != 0;
    for (int i = threadIdx.x; i < sz; i += blockDim.x) {
        int id = 0x7fffffff;
        if (i < total) {
            if (all_items) id = i;
            else if (i < c0) id = cand[(size_t)ul * cmax + i];
            else if (i < c0 + c1) id = cand[(size_t)ul * cmax + cmax / 2 + (i - c0)];
            else id = nan_list[i - c0 - c1];
        }
        s_id[i] = id;
        s_sc[i] = i < total ? b : __int_as_float(0xff800000);  // out-of-range rows and NaN users score b
    }
    __syncthreads();
    // a NaN user row makes every score b: the exact answer is items 0..topk-1, whatever the candidates were
    if (!all_items && u_ok && u_nan) {
        for (int j = threadIdx.x; j < topk; j += blockDim.x) {
            idx_out[(size_t)(user0 + ul) * topk + j] = j < n ? j : -1;
            if (score_out) score_out[(size_t)(user0 + ul) * topk + j] = j < n ? b : 0.f;
        }
        return;
    }
    if (u_ok)
        for (int base = 0; base < total; base += rows) {
            const int cnt = min(rows, total - base);
            for (int r = warp; r < cnt; r += nwarps) {  // coalesced: a warp per row, all rows of the chunk in flight
                const int id = s_id[base + r];
                const bool ok = id >= 0 && id < n;
                for (int d = lane; d < k; d += 32) {
                    float *dst = s_q + (size_t)r * stride + d;
                    if (ok)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(Q + (size_t)id * k + d)
                                     : "memory");
                    else
                        *dst = 0.f;
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
                const int id = s_id[base + r];
                if (id >= 0 && id < n) {
                    const float *q = s_q + (size_t)r * stride;
                    float z = 0.0f;
                    for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(s_p[d], q[d]));
                    s_sc[base + r] = isnan(z) ? b : z;  // mf/mf.cpp:4305-4306
                }
            }
            __syncthreads();
        }
    for (int size = 2; size <= sz; size <<= 1)
        for (int stride2 = size >> 1; stride2 > 0; stride2 >>= 1) {
            for (int i = threadIdx.x; i < sz / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride2 - 1)), hi = lo + stride2;
                const bool up = (lo & size) == 0;  // ascending block: "before" first
                const float a_s = s_sc[lo], b_s = s_sc[hi];
                const int a_i = s_id[lo], b_i = s_id[hi];
                const bool swap = up ? before(b_s, b_i, a_s, a_i) : before(a_s, a_i, b_s, b_i);
                if (swap) {
                    s_sc[lo] = b_s; s_sc[hi] = a_s;
                    s_id[lo] = b_i; s_id[hi] = a_i;
                }
            }
            __syncthreads();
        }
    const int valid = min(total, n);
    for (int j = threadIdx.x; j < topk; j += blockDim.x) {
        const bool ok = j < valid && j < sz && s_id[j] != 0x7fffffff;
        idx_out[(size_t)(user0 + ul) * topk + j] = ok ? s_id[j] : -1;
        if (score_out) score_out[(size_t)(user0 + ul) * topk + j] = ok ? s_sc[j] : 0.f;
    }
}

// ---- host helpers -----------------------------------------------------------------------------------------------
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
    return 1024 + (size_t)KATOMS * TK_M * 128 + (size_t)TK_STAGES * KATOMS * TK_N * 128 + 4 * TK_N * sizeof(float) + 16 * 8;
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

int launch_select(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers, int user0,
                  const int *cand, const int *cand_cnt, int cmax, int ub, const int *nan_list, const int *nan_count,
                  int all_items, int topk, int *idx_out, float *score_out, int *overflow, cudaStream_t st) {
    const int stride = k | 1;
    int rows = (48 * 1024 - ((k + 3) & ~3) * 4) / (stride * 4);  // ~48 KB of staging per block
    rows = rows > 256 ? 256 : rows;
    if (rows < 1) return (int)cudaErrorNotSupported;
    const size_t smem = (size_t)(((k + 3) & ~3) + (size_t)rows * stride) * 4;
    cudaError_t e = cudaFuncSetAttribute(k_topk_select<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_topk_select<2048><<<nusers, 256, smem, st>>>(P, Q, m, n, k, b, users, nusers, user0, cand, cand_cnt, cmax, ub, nan_list,
                                                    nan_count, all_items, topk, rows, idx_out, score_out, overflow);
    return (int)cudaGetLastError();
}

}  // namespace

// ================================================================================================================
extern "C" {

int mfk_topk_max_candidates(void) { return 1920; }

// Everything on `stream`; P [m][k], Q [n][k] fp32 on the device; users [nusers] on the device; outputs on the device.
// work: caller-provided device scratch of mfk_topk_work_bytes(...) bytes.  *overflow_dev is set to 1 if a candidate
// list overflowed (the result of that user is then not guaranteed; the caller fails loudly).
static int topk_gpt(int n_samp, int topk) { return n_samp < 4 * topk ? 8 : 2; }  // maxima per tile: per 32-column chunk or per half tile

size_t mfk_topk_work_bytes(int n, int k, int batch_users, int sample_stride) {
    const int kp = ((k + TK_KATOM - 1) / TK_KATOM) * TK_KATOM;
    const size_t npad = ((size_t)n + TK_N - 1) / TK_N * TK_N;
    const size_t ub = ((size_t)batch_users + TK_M - 1) / TK_M * TK_M;
    const size_t n_tiles = npad / TK_N, n_samp = (n_tiles + sample_stride - 1) / sample_stride;
    size_t b = 0;
    b += npad * kp * 2 + 4 * npad * 4 + 256;                     // Q bf16, norm, is_nan, qn_max, qn_cand
    b += ub * kp * 2 + 4 * ub * 4 + 256;                          // P bf16, norm, is_nan, eps, tau
    b += n_samp * 8 * ub * 4 + ub * (size_t)mfk_topk_max_candidates() * 4 + 2 * ub * 4;  // maxes, cand, cand_cnt
    b += 1024 * 4 + 64;                                          // nan list, counters
    return b + 4096;
}

int mfk_topk(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers, int topk,
             int *idx_out, float *score_out, void *work, size_t work_bytes, int batch_users, int sample_stride,
             int sm_count, int *overflow_dev, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nusers <= 0) return 0;
    const int cmax = mfk_topk_max_candidates();
    // small item sets: every item is a candidate, no GEMM
    if (n + 0 <= 2048 && topk <= 2048) {
        return launch_select(P, Q, m, n, k, b, users, nusers, 0, nullptr, nullptr, 0, 0, nullptr, nullptr, 1, topk, idx_out,
                             score_out, overflow_dev, st);
    }
    const int kp = ((k + TK_KATOM - 1) / TK_KATOM) * TK_KATOM;
    if (kp > 128 || topk > 128 || topk < 1) return (int)cudaErrorNotSupported;
    if (work_bytes < mfk_topk_work_bytes(n, k, batch_users, sample_stride)) return (int)cudaErrorInvalidValue;
    const int npad = (n + TK_N - 1) / TK_N * TK_N, ub = (batch_users + TK_M - 1) / TK_M * TK_M;
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
    float *qn_max = (float *)take((size_t)npad * 4), *qn_cand = (float *)take((size_t)npad * 4);
    __nv_bfloat16 *Pb = (__nv_bfloat16 *)take((size_t)ub * kp * 2);
    float *pnorm = (float *)take((size_t)ub * 4);
    int *p_nan = (int *)take((size_t)ub * 4);
    float *eps = (float *)take((size_t)ub * 4), *tau = (float *)take((size_t)ub * 4);
    const int gpt = topk_gpt(n_samp, topk);
    float *maxes = (float *)take((size_t)n_samp * gpt * ub * 4);
    int *cand = (int *)take((size_t)ub * cmax * 4), *cand_cnt = (int *)take((size_t)2 * ub * 4);
    int *nan_list = (int *)take(1024 * 4), *nan_count = (int *)take(64);

    k_topk_prep<<<148 * 8, 256, 0, st>>>(Q, n, k, nullptr, n, npad, kp, Qb, qnorm, q_nan);
    k_topk_item_bounds<<<(npad + 255) / 256, 256, 0, st>>>(qnorm, q_nan, npad, qn_max, qn_cand);
    k_topk_nan_list<<<1, 1024, 0, st>>>(q_nan, n, topk, nan_list, nan_count);
    CUtensorMap tmQ, tmP;
    if (make_tmap(&tmQ, Qb, npad, kp, TK_N) || make_tmap(&tmP, Pb, ub, kp, TK_M)) return (int)cudaErrorUnknown;

    for (int u0 = 0; u0 < nusers; u0 += batch_users) {
        const int nu = nusers - u0 < batch_users ? nusers - u0 : batch_users;
        const int nu_pad = (nu + TK_M - 1) / TK_M * TK_M;
        k_topk_prep<<<148 * 4, 256, 0, st>>>(P, m, k, users + u0, nu, ub, kp, Pb, pnorm, p_nan);
        k_topk_user_eps<<<(ub + 255) / 256, 256, 0, st>>>(pnorm, p_nan, ub, eps);
        TopkGemmArgs a;
        a.n_user_tiles = nu_pad / TK_M;
        a.n_item_tiles = n_tiles;
        a.ub = ub;
        a.eps = eps;
        a.tau = tau;
        a.cand = cand;
        a.cand_cnt = cand_cnt;
        a.cmax = cmax;
        a.maxes = maxes;
        // pass A: tile maxima of the lower bound on every sample_stride-th tile
        a.tile_stride = sample_stride;
        a.gpt = gpt;
        a.qn = qn_max;
        int rc = kp == 64 ? launch_gemm<1, MODE_MAX>(tmP, tmQ, a, sm_count, st) : launch_gemm<2, MODE_MAX>(tmP, tmQ, a, sm_count, st);
        if (rc) return rc;
        if (n_samp * gpt <= 32 * 16)
            k_topk_tau<16><<<(ub * 32 + 255) / 256, 256, 0, st>>>(maxes, n_samp * gpt, ub, nu, p_nan, topk, tau);
        else
            k_topk_tau<64><<<(ub * 32 + 255) / 256, 256, 0, st>>>(maxes, n_samp * gpt, ub, nu, p_nan, topk, tau);
        // pass C: candidates
        a.tile_stride = 1;
        a.qn = qn_cand;
        rc = kp == 64 ? launch_gemm<1, MODE_CAND>(tmP, tmQ, a, sm_count, st) : launch_gemm<2, MODE_CAND>(tmP, tmQ, a, sm_count, st);
        if (rc) return rc;
        rc = launch_select(P, Q, m, n, k, b, users, nu, u0, cand, cand_cnt, cmax, ub, nan_list, nan_count, 0, topk, idx_out,
                           score_out, overflow_dev, st);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
