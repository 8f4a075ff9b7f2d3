// csrc/kernels.cu -- sm_100a kernels of the matrix-factorisation hot path + their launchers.
//
// Kernels (reference lines they replace are cited at each one; paths relative to /root/reference):
//   k_sgd_ring_epoch   throughput SGD: one warp per rating update, float4 row accesses, warp-shuffle
//                      dot product fused with the regularised AdaGrad step, conflict-free two-level
//                      ring schedule (CTA ring over column bands, warp ring over sub-bands)
//   k_sgd_exact_level  bit-exact SGD in the reference's sequential order, one wavefront per launch
//   k_stats / k_ring_keys / k_ring_gather / k_init_rows / k_finalize_rows   preprocessing of fpsg
//   k_predict_pairs / k_sq_err / k_reg2                                     predict + metrics
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -ftz=true (the reference runs its loop with
// flush-to-zero on, mf/mf.cpp:2788-2791).  Exact kernels use __f*_rn intrinsics so that nothing
// is contracted into an FMA (the reference is built without FMA, mf/CMakeLists.txt:10).

#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "rsqrt12_table.h"

namespace {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cta_smem(const unsigned *p) {
    unsigned v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_smem(unsigned *p, unsigned v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_add_acqrel_cta_smem(unsigned *p, unsigned v) {
    unsigned old;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ double block_sum_double(double v, double *smem /* >= 32 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (int)((blockDim.x + 31) >> 5) ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    }
    return v;  // valid in warp 0
}

// The reference's eta * _mm_rsqrt_ps(G) (mf/mf.cpp:1469-1470): exact emulation of the x86 12-bit
// approximation through the verified table (oracle/gen_rsqrt_table.c).  G >= 1 in practice.
__device__ __forceinline__ float rsqrt12(float x, const unsigned *__restrict__ tab) {
    const unsigned b = __float_as_uint(x);
    const unsigned e = (b >> 23) & 0xffu, man = b & 0x7fffffu;
    unsigned out;
    if (e == 0xffu)
        out = man ? (b | 0x400000u) : ((b >> 31) ? 0xffc00000u : 0u);
    else if (e == 0u)
        out = (b & 0x80000000u) | 0x7f800000u;
    else if (b >> 31)
        out = 0xffc00000u;
    else {
        const unsigned p = (e & 1u) ? 0u : 1u;
        const int sh = ((int)e - (127 + (int)p)) / 2;
        out = __ldg(tab + p * 1024u + (man >> 13)) - ((unsigned)sh << 23);
    }
    return __uint_as_float(out);
}

__device__ unsigned g_rsqrt12_table[2048];

// ------------------------------------------------------------------------------------------------
// collect_info (mf/mf.cpp:462-484)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stats(const mfk_node *__restrict__ R, long long nnz, double *out2) {
    __shared__ double sm[2][32];
    double s = 0.0, s2 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
         i += (long long)gridDim.x * blockDim.x) {
        const double r = (double)R[i].r;
        s += r;
        s2 += r * r;
    }
    s = block_sum_double(s, sm[0]);
    __syncthreads();
    s2 = block_sum_double(s2, sm[1]);
    if (threadIdx.x == 0) {
        atomicAdd(out2, s);
        atomicAdd(out2 + 1, s2);
    }
}

// ------------------------------------------------------------------------------------------------
// ring preprocessing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ring_sub_block(const mfk_ring_shape &sh, int a, int b) {
    const int c = a / sh.segA1, w = (a - c * sh.segA1) / sh.segA2;
    const int j = b / sh.segB1, i = (b - j * sh.segB1) / sh.segB2;
    return (((unsigned long long)c * sh.nB1 + j) * sh.nW + w) * sh.nB2 + i;
}

__global__ void __launch_bounds__(256)
k_ring_keys(const mfk_node *__restrict__ R, long long nnz, const int *__restrict__ p_map,
            const int *__restrict__ q_map, int swap_sides, mfk_ring_shape sh, int *omega_p, int *omega_q,
            unsigned long long *keys, unsigned *vals) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
         i += (long long)gridDim.x * blockDim.x) {
        const mfk_node N = R[i];
        const int u = p_map[N.u], v = q_map[N.v];
        atomicAdd(omega_p + u, 1);
        atomicAdd(omega_q + v, 1);
        const int a = swap_sides ? v : u, b = swap_sides ? u : v;
        keys[i] = (ring_sub_block(sh, a, b) << sh.bitsA) | (unsigned long long)a;
        vals[i] = (unsigned)i;
    }
}

__global__ void __launch_bounds__(256)
k_ring_gather(const mfk_node *__restrict__ R, long long nnz, const unsigned long long *__restrict__ keys,
              const unsigned *__restrict__ vals, const int *__restrict__ p_map,
              const int *__restrict__ q_map, int swap_sides, mfk_ring_shape sh, float inv_scale, int *ra,
              int *rb, float *rr, unsigned *sub_off) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
         i += (long long)gridDim.x * blockDim.x) {
        const mfk_node N = R[vals[i]];
        const int u = p_map[N.u], v = q_map[N.v];
        ra[i] = swap_sides ? v : u;
        rb[i] = swap_sides ? u : v;
        rr[i] = inv_scale == 1.0f ? N.r : N.r * inv_scale;  // scale_problem, mf/mf.cpp:517-527
        const long long sb = (long long)(keys[i] >> sh.bitsA);
        const long long prev = i > 0 ? (long long)(keys[i - 1] >> sh.bitsA) : -1;
        for (long long s = prev + 1; s <= sb; s++) sub_off[s] = (unsigned)i;
        if (i == nnz - 1)
            for (long long s = sb + 1; s <= sh.nSub; s++) sub_off[s] = (unsigned)nnz;
    }
}

// ------------------------------------------------------------------------------------------------
// init_model (mf/mf.cpp:952-1007) with jump-ahead in minstd_rand0
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mulmod31(unsigned a, unsigned b) {
    return (unsigned)(((unsigned long long)a * b) % 2147483647ull);
}
__device__ unsigned powmod31(unsigned base, unsigned long long e) {
    unsigned r = 1;
    while (e) {
        if (e & 1ull) r = mulmod31(r, base);
        base = mulmod31(base, base);
        e >>= 1;
    }
    return r;
}

__global__ void k_flag_nonempty(const int *__restrict__ omega, int rows, int *flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) flags[i] = omega[i] > 0 ? 1 : 0;
}
__global__ void k_rank_total(const int *rank, const int *omega, int rows, int *total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *total = rows > 0 ? rank[rows - 1] + (omega[rows - 1] > 0 ? 1 : 0) : 0;
}

// one thread per (row, 8-dim chunk).  Draw number t (1-based) of the engine is 16807^t mod (2^31-1);
// the distribution returns float(x-1)/2^31 (clamped below 1); the factor is sqrt(1/k) as float.
__global__ void __launch_bounds__(256)
k_init_rows(float *M, float *G, const int *__restrict__ omega, const int *__restrict__ rank, int rank_base,
            int rows, int k, int k_al, float s) {
    const int chunks = k_al >> 3;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tid >= (long long)rows * chunks) return;
    const int row = (int)(tid / chunks), ch = (int)(tid - (long long)row * chunks);
    float *dst = M + (size_t)row * k_al + ch * 8;
    if (ch == 0) {
        G[2 * (size_t)row] = 1.0f;  // PG, QG = 1: mf/mf.cpp:2835
        G[2 * (size_t)row + 1] = 1.0f;
    }
    const bool seen = omega[row] > 0;
    unsigned x = 0;
    if (seen && ch * 8 < k)
        x = powmod31(16807u, (unsigned long long)(rank[row] + rank_base) * k + ch * 8 + 1);
    float out[8];
#pragma unroll
    for (int d = 0; d < 8; d++) {
        const int dim = ch * 8 + d;
        float val = 0.0f;
        if (dim < k) {
            if (seen) {
                float f = (float)(x - 1u) * (1.0f / 2147483648.0f);
                if (f >= 1.0f) f = 0.99999994f;
                val = __fmul_rn(f, s);
                x = mulmod31(x, 16807u);
            } else {
                val = __uint_as_float(0x7fc00000u);  // quiet NaN, mf/mf.cpp:996-999
            }
        }
        out[d] = val;
    }
    reinterpret_cast<float4 *>(dst)[0] = make_float4(out[0], out[1], out[2], out[3]);
    reinterpret_cast<float4 *>(dst)[1] = make_float4(out[4], out[5], out[6], out[7]);
}

// ------------------------------------------------------------------------------------------------
// The throughput kernel.
//
// Replaces the per-rating loop SolverBase::run + L2_MFR::prepare_for_sg_update + MFSolver::sg_update
// (mf/mf.cpp:1220-1235, 1720-1728, 1462-1548) and the block scheduler (mf/mf.cpp:113-150,193-220).
//
// Schedule (race-free by construction, like the reference's scheduler that never co-schedules two
// blocks sharing a row band or a column band, mf/mf.cpp:130-142):
//   * CTA c owns row band c for the whole run; warp w of the CTA owns sub-row-band w.  Rows of the
//     owned side are therefore only ever touched by one warp: no synchronisation, and a row stays in
//     registers across a run of consecutive ratings of the same row.
//   * Column bands rotate ring-wise: at global step g, CTA c works on band (c*S1 + g) mod nB1.  The
//     previous user of that band is CTA c+1 at step g-S1, so CTA c only waits for its neighbour's
//     progress counter (acquire/release in global memory).  No grid-wide barrier.
//   * Inside a step the same ring runs one level down: at sub-step t2 warp w works on sub-band
//     (w*S2 + t2) mod nB2 of the band and waits only for warp w+1's counter in shared memory.
//   The order of updates of every row is fixed by the schedule, so a run is reproducible.
//
// Arithmetic per rating: lane l holds dims 4l..4l+3 of both rows (one 128-bit
// load each); z by butterfly shuffle; e = r - z; g_p = lambda_p p - e q, g_q = lambda_q q - e p from
// the OLD p,q; p -= eta rsqrt(G_p) g_p; G += sum(g^2)/8 for BOTH halves (the shipped SSE path's rk,
// SURVEY.md F2); dims 0-7 and 8..k_al have separate accumulators; epoch 0 touches dims 0-7 only.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(1024, 1) k_sgd_ring_epoch(const __grid_constant__ mfk_ring_args g) {
    __shared__ unsigned s_wprog[kWarp];
    __shared__ unsigned s_done;

    const mfk_ring_shape &sh = g.shape;
    const int c = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = g.k_al >> 2;
    const bool full = g.epoch > 0;  // slow_only == false

    if (threadIdx.x < kWarp) s_wprog[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();

    bool act[NV], h0[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) {
        act[j] = lane + 32 * j < nvec;
        h0[j] = lane + 32 * j < 2;
    }

    float4 p[NV], q[NV];
    float ag0 = 1.f, ag1 = 1.f;  // AdaGrad accumulators of the row held in registers
    int cur_a = -1;
    double loss = 0.0;
    bool dead = false;  // a wait timed out: stop waiting so that the kernel still terminates

    auto flush_row = [&]() {
        if (cur_a < 0) return;
        float *row = g.A + (size_t)cur_a * g.k_al;
#pragma unroll
        for (int j = 0; j < NV; j++)
            if (act[j] && (full || h0[j])) reinterpret_cast<float4 *>(row)[lane + 32 * j] = p[j];
        if (lane == 0) reinterpret_cast<float2 *>(g.AG)[cur_a] = make_float2(ag0, ag1);
    };

    for (int t = 0; t < sh.nB1; ++t) {
        const unsigned gstep = (unsigned)g.epoch * (unsigned)sh.nB1 + (unsigned)t;
        const int band = (int)(((unsigned)c * (unsigned)sh.S1 + gstep) % (unsigned)sh.nB1);

        // ---- acquire the column band: wait until the neighbour CTA has released it ----
        if (sh.nC > 1 && gstep >= (unsigned)sh.S1 && !dead) {
            const unsigned need = gstep - (unsigned)sh.S1 + 1u;
            if (lane == 0) {
                const unsigned *flag = g.progress + (c + 1) % sh.nC;
                unsigned spins = 0;
                while (ld_acquire_gpu(flag) < need) {
                    __nanosleep(64);
                    if (++spins > (1u << 24)) {
                        atomicExch(g.error_flag, 1);
                        break;
                    }
                }
            }
            __syncwarp();
            dead = __shfl_sync(kFull, (int)(*(volatile int *)g.error_flag != 0), 0) != 0;
        }

        const long long sub_base = (((long long)c * sh.nB1 + band) * sh.nW + w) * sh.nB2;
        const unsigned my_off = lane <= sh.nB2 ? g.sub_off[sub_base + lane] : 0u;

        for (int t2 = 0; t2 < sh.nB2; ++t2) {
            const unsigned hstep = (unsigned)t * (unsigned)sh.nB2 + (unsigned)t2;
            const int sub = (w * sh.S2 + t2) % sh.nB2;

            // ---- acquire the sub-band from the neighbour warp ----
            if (sh.nW > 1 && t2 >= sh.S2 && !dead) {
                const unsigned need = hstep - (unsigned)sh.S2 + 1u;
                if (lane == 0) {
                    const unsigned *flag = &s_wprog[(w + 1) % sh.nW];
                    unsigned spins = 0;
                    while (ld_acquire_cta_smem(flag) < need) {
                        if (++spins > (1u << 27)) {
                            atomicExch(g.error_flag, 2);
                            break;
                        }
                    }
                }
                __syncwarp();
            }

            const unsigned beg = __shfl_sync(kFull, my_off, sub);
            const unsigned end = __shfl_sync(kFull, my_off, sub + 1);

            for (unsigned base = beg; base < end; base += 32) {
                const unsigned mine = base + lane;
                const int my_a = mine < end ? g.ra[mine] : -1;
                const int my_b = mine < end ? g.rb[mine] : -1;
                const float my_r = mine < end ? g.rr[mine] : 0.f;
                const int cnt = min(32u, end - base);

                for (int i = 0; i < cnt; ++i) {
                    const int a = __shfl_sync(kFull, my_a, i);
                    const int b = __shfl_sync(kFull, my_b, i);
                    const float r = __shfl_sync(kFull, my_r, i);

                    if (a != cur_a) {  // warp-uniform: a new run of the owned row
                        flush_row();
                        cur_a = a;
                        const float *row = g.A + (size_t)a * g.k_al;
#pragma unroll
                        for (int j = 0; j < NV; j++)
                            p[j] = act[j] ? reinterpret_cast<const float4 *>(row)[lane + 32 * j]
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float2 ag = reinterpret_cast<const float2 *>(g.AG)[a];
                        ag0 = ag.x;
                        ag1 = ag.y;
                    }
                    float *qrow = g.B + (size_t)b * g.k_al;
#pragma unroll
                    for (int j = 0; j < NV; j++)
                        q[j] = act[j] ? reinterpret_cast<const float4 *>(qrow)[lane + 32 * j]
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    float2 bg = reinterpret_cast<const float2 *>(g.BG)[b];

                    // z = <p,q>  (calc_z, mf/mf.cpp:1264-1273)
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; j++)
                        part += p[j].x * q[j].x + p[j].y * q[j].y + p[j].z * q[j].z + p[j].w * q[j].w;
                    const float e = r - warp_sum(part);  // mf/mf.cpp:1724
                    if (lane == 0) loss += (double)(e * e);  // mf/mf.cpp:1725-1726

                    // sg_update for both halves (mf/mf.cpp:1462-1548, 1228-1234)
                    const float eta_p0 = g.eta * rsqrtf(ag0), eta_q0 = g.eta * rsqrtf(bg.x);
                    const float eta_p1 = g.eta * rsqrtf(ag1), eta_q1 = g.eta * rsqrtf(bg.y);
                    float sp0 = 0.f, sq0 = 0.f, sp1 = 0.f, sq1 = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; j++) {
                        if (!(full || h0[j])) continue;
                        const float ep = h0[j] ? eta_p0 : eta_p1, eq = h0[j] ? eta_q0 : eta_q1;
                        float sp = 0.f, sq = 0.f;
#define MFB_UPD(X)                                             \
    {                                                          \
        const float gp = g.lambda_a * p[j].X - e * q[j].X;     \
        const float gq = g.lambda_b * q[j].X - e * p[j].X;     \
        sp += gp * gp;                                         \
        sq += gq * gq;                                         \
        p[j].X -= ep * gp;                                     \
        q[j].X -= eq * gq;                                     \
    }
                        MFB_UPD(x) MFB_UPD(y) MFB_UPD(z) MFB_UPD(w)
#undef MFB_UPD
                        if (h0[j]) {
                            sp0 += sp;
                            sq0 += sq;
                        } else {
                            sp1 += sp;
                            sq1 += sq;
                        }
                    }
                    // half 0 lives in lanes 0,1 (vector 0); half 1 everywhere else
                    sp0 += __shfl_xor_sync(kFull, sp0, 1);
                    sq0 += __shfl_xor_sync(kFull, sq0, 1);
                    sp0 = __shfl_sync(kFull, sp0, 0);
                    sq0 = __shfl_sync(kFull, sq0, 0);
                    ag0 += sp0 * 0.125f;
                    bg.x += sq0 * 0.125f;
                    if (full) {
                        ag1 += warp_sum(sp1) * 0.125f;  // rk_slow for both halves: SURVEY.md F2
                        bg.y += warp_sum(sq1) * 0.125f;
                    }
#pragma unroll
                    for (int j = 0; j < NV; j++)
                        if (act[j] && (full || h0[j])) reinterpret_cast<float4 *>(qrow)[lane + 32 * j] = q[j];
                    if (lane == 0) reinterpret_cast<float2 *>(g.BG)[b] = bg;
                    __syncwarp();  // lane 0's accumulator store is re-read by every lane if b repeats
                }
            }

            // ---- release the sub-band to the next warp of the ring ----
            __syncwarp();
            if (lane == 0) st_release_cta_smem(&s_wprog[w], hstep + 1u);
        }

        // ---- this warp is done with the band; the last warp of the CTA publishes the step ----
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            const unsigned old = atom_add_acqrel_cta_smem(&s_done, 1u);
            if (old + 1u == (unsigned)sh.nW * (unsigned)(t + 1)) {
                __threadfence();
                st_release_gpu(g.progress + c, gstep + 1u);
            }
        }
        __syncwarp();
    }

    flush_row();
    if (lane == 0 && loss != 0.0) atomicAdd(g.loss, loss);
}

// ------------------------------------------------------------------------------------------------
// The exact kernel: one thread per rating of one wavefront level; arithmetic is the SSE code path's,
// operation for operation, so results equal the reference's bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void exact_half(float *p, float *q, float *pG, float *qG, float e, int d0, int d1,
                                           float lp, float lq, float eta, const unsigned *tab) {
    const float eta_p = __fmul_rn(eta, rsqrt12(*pG, tab));
    const float eta_q = __fmul_rn(eta, rsqrt12(*qG, tab));
    float sp[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = d0; d < d1; d += 4) {
        float4 pv = *reinterpret_cast<float4 *>(p + d), qv = *reinterpret_cast<float4 *>(q + d);
        float *pp = reinterpret_cast<float *>(&pv), *qq = reinterpret_cast<float *>(&qv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float gp = __fsub_rn(__fmul_rn(lp, pp[j]), __fmul_rn(e, qq[j]));
            const float gq = __fsub_rn(__fmul_rn(lq, qq[j]), __fmul_rn(e, pp[j]));
            sp[j] = __fadd_rn(sp[j], __fmul_rn(gp, gp));
            sq[j] = __fadd_rn(sq[j], __fmul_rn(gq, gq));
            pp[j] = __fsub_rn(pp[j], __fmul_rn(eta_p, gp));
            qq[j] = __fsub_rn(qq[j], __fmul_rn(eta_q, gq));
        }
        *reinterpret_cast<float4 *>(p + d) = pv;
        *reinterpret_cast<float4 *>(q + d) = qv;
    }
    *pG = __fadd_rn(*pG, __fmul_rn(__fadd_rn(__fadd_rn(sp[0], sp[1]), __fadd_rn(sp[2], sp[3])), 0.125f));
    *qG = __fadd_rn(*qG, __fmul_rn(__fadd_rn(__fadd_rn(sq[0], sq[1]), __fadd_rn(sq[2], sq[3])), 0.125f));
}

__global__ void __launch_bounds__(128)
k_sgd_exact_level(const mfk_node *__restrict__ R, const unsigned *__restrict__ order, int count, float *P,
                  float *Q, float *PG, float *QG, int k_al, float lp, float lq, float eta, int slow_only,
                  float *e2_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned idx = order[i];
    const mfk_node N = R[idx];
    float *p = P + (size_t)N.u * k_al, *q = Q + (size_t)N.v * k_al;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int d = 0; d < k_al; d += 4) {
        const float4 a = *reinterpret_cast<const float4 *>(p + d), b = *reinterpret_cast<const float4 *>(q + d);
        l0 = __fadd_rn(l0, __fmul_rn(a.x, b.x));
        l1 = __fadd_rn(l1, __fmul_rn(a.y, b.y));
        l2 = __fadd_rn(l2, __fmul_rn(a.z, b.z));
        l3 = __fadd_rn(l3, __fmul_rn(a.w, b.w));
    }
    const float z = __fadd_rn(__fadd_rn(l0, l1), __fadd_rn(l2, l3));
    const float e = __fsub_rn(N.r, z);
    e2_out[idx] = __fmul_rn(e, e);
    exact_half(p, q, PG + 2 * (size_t)N.u, QG + 2 * (size_t)N.v, e, 0, 8, lp, lq, eta, g_rsqrt12_table);
    if (!slow_only)
        exact_half(p, q, PG + 2 * (size_t)N.u + 1, QG + 2 * (size_t)N.v + 1, e, 8, k_al, lp, lq, eta,
                   g_rsqrt12_table);
}

__global__ void __launch_bounds__(256) k_sum_f32(const float *__restrict__ x, long long n, double *out) {
    __shared__ double sm[32];
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        s += (double)x[i];
    s = block_sum_double(s, sm);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// calc_reg2's inner sum (mf/mf.cpp:608-633): sum_i omega_i * <row_i,row_i>, the inner product in the
// SSE lane order (557-566), int*float product in float, accumulation in double.
__global__ void __launch_bounds__(128)
k_reg2(const float *__restrict__ M, const int *__restrict__ omega, int rows, int k_al, double *out) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        const int om = omega[i];
        if (om <= 0) continue;
        const float *row = M + (size_t)i * k_al;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        for (int d = 0; d < k_al; d += 4) {
            const float4 a = *reinterpret_cast<const float4 *>(row + d);
            l0 = __fadd_rn(l0, __fmul_rn(a.x, a.x));
            l1 = __fadd_rn(l1, __fmul_rn(a.y, a.y));
            l2 = __fadd_rn(l2, __fmul_rn(a.z, a.z));
            l3 = __fadd_rn(l3, __fmul_rn(a.w, a.w));
        }
        acc += (double)__fmul_rn((float)om, __fadd_rn(__fadd_rn(l0, l1), __fadd_rn(l2, l3)));
    }
    acc = block_sum_double(acc, sm);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// scale_model + shrink_model + shuffle_model in one pass (mf/mf.cpp:529-553,1057-1074,1027-1055)
__global__ void __launch_bounds__(256)
k_finalize_rows(const float *__restrict__ M, const int *__restrict__ map, int rows, int k, int k_al, float factor,
                float *out) {
    const long long total = (long long)rows * k;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int id = (int)(t / k), d = (int)(t - (long long)id * k);
        const float v = M[(size_t)map[id] * k_al + d];
        // NaN rows keep their bit pattern (x86 returns the quiet-NaN operand; the GPU would canonicalise it)
        out[t] = (factor == 1.0f || isnan(v)) ? v : __fmul_rn(v, factor);
    }
}

// mf_predict (mf/mf.cpp:4295-4314): bounds -> b; z = sum in index order starting from 0.0f with the
// product rounded before the add; NaN -> b.
__device__ __forceinline__ float predict_exact(const float *__restrict__ P, const float *__restrict__ Q, int m,
                                               int n, int k, float b, int u, int v) {
    if (u < 0 || u >= m || v < 0 || v >= n) return b;
    const float *p = P + (size_t)u * k, *q = Q + (size_t)v * k;
    float z = 0.0f;
    for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(p[d], q[d]));
    return isnan(z) ? b : z;
}

__global__ void __launch_bounds__(256)
k_predict_pairs(const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k, float b,
                const float *__restrict__ pairs, long long npairs, float *out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npairs;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = predict_exact(P, Q, m, n, k, b, (int)pairs[2 * i], (int)pairs[2 * i + 1]);
}

__global__ void __launch_bounds__(256)
k_sq_err(const mfk_node *__restrict__ R, long long nnz, const float *__restrict__ P, const float *__restrict__ Q,
         int m, int n, int k, float b, double *out) {
    __shared__ double sm[32];
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
         i += (long long)gridDim.x * blockDim.x) {
        const mfk_node N = R[i];
        const float e = __fsub_rn(N.r, predict_exact(P, Q, m, n, k, b, N.u, N.v));
        s += (double)__fmul_rn(e, e);
    }
    s = block_sum_double(s, sm);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

inline int grid_for(long long n, int block, int cap) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

bool g_table_ready[64] = {false};

int ensure_table() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 64 && g_table_ready[dev]) return 0;
    e = cudaMemcpyToSymbol(g_rsqrt12_table, MFB200_RSQRT12_TABLE, sizeof(unsigned) * 2048);
    if (e != cudaSuccess) return (int)e;
    if (dev < 64) g_table_ready[dev] = true;
    return 0;
}

}  // namespace

// ================================================================================================
// launchers (C ABI)
// ================================================================================================
extern "C" {

int mfk_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    return n;
}

int mfk_stats(const mfk_node *R, long long nnz, double *out2, void *stream) {
    k_stats<<<grid_for(nnz, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(R, nnz, out2);
    return (int)cudaGetLastError();
}

int mfk_ring_keys(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, int swap_sides,
                  mfk_ring_shape shape, int *omega_p, int *omega_q, unsigned long long *keys, unsigned *vals,
                  void *stream) {
    k_ring_keys<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        R, nnz, p_map, q_map, swap_sides, shape, omega_p, omega_q, keys, vals);
    return (int)cudaGetLastError();
}

size_t mfk_sort_tmp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr,
                                    (unsigned long long *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, n, 0, 64);
    return bytes;
}

int mfk_sort_pairs(unsigned long long *keys_in, unsigned long long *keys_out, unsigned *vals_in,
                   unsigned *vals_out, long long n, int end_bit, void *tmp, size_t tmp_bytes, void *stream) {
    return (int)cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned long long *)keys_in, keys_out,
                                                (const unsigned *)vals_in, vals_out, n, 0, end_bit,
                                                (cudaStream_t)stream);
}

int mfk_ring_gather(const mfk_node *R, long long nnz, const unsigned long long *keys_sorted,
                    const unsigned *vals_sorted, const int *p_map, const int *q_map, int swap_sides,
                    mfk_ring_shape shape, float inv_scale, int *ra, int *rb, float *rr, unsigned *sub_off,
                    void *stream) {
    k_ring_gather<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        R, nnz, keys_sorted, vals_sorted, p_map, q_map, swap_sides, shape, inv_scale, ra, rb, rr, sub_off);
    return (int)cudaGetLastError();
}

size_t mfk_rank_tmp_bytes(int rows) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, rows);
    return bytes + sizeof(int) * (size_t)(rows > 0 ? rows : 1) + 256;
}

int mfk_exclusive_rank(const int *omega, int rows, int *rank, int *total_out_dev, void *tmp, size_t tmp_bytes,
                       void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    // tmp layout: [flags: rows ints, padded to 256 B][cub scratch]
    int *flags = (int *)tmp;
    size_t off = ((sizeof(int) * (size_t)(rows > 0 ? rows : 1)) + 255) & ~(size_t)255;
    if (rows > 0) {
        k_flag_nonempty<<<(rows + 255) / 256, 256, 0, st>>>(omega, rows, flags);
        size_t cub_bytes = tmp_bytes - off;
        cudaError_t e = cub::DeviceScan::ExclusiveSum((char *)tmp + off, cub_bytes, (const int *)flags, rank, rows, st);
        if (e != cudaSuccess) return (int)e;
    }
    k_rank_total<<<1, 32, 0, st>>>(rank, omega, rows, total_out_dev);
    return (int)cudaGetLastError();
}

int mfk_init_rows(float *M, float *G, const int *omega, const int *rank, int rank_base, int rows, int k, int k_al,
                  void *stream) {
    const long long threads = (long long)rows * (k_al / 8);
    if (threads == 0) return 0;
    const float s = (float)sqrt(1.0 / (double)k);  // mf/mf.cpp:971
    k_init_rows<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, G, omega, rank, rank_base,
                                                                                    rows, k, k_al, s);
    return (int)cudaGetLastError();
}

int mfk_sgd_ring_epoch(const mfk_ring_args *args, void *stream) {
    const int nvec = args->k_al / 4;
    const int nv = (nvec + 31) / 32;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block(args->shape.nW * 32);
    const void *fn = nullptr;
    if (nv <= 1)
        fn = (const void *)k_sgd_ring_epoch<1>;
    else if (nv == 2)
        fn = (const void *)k_sgd_ring_epoch<2>;
    else if (nv <= 4)
        fn = (const void *)k_sgd_ring_epoch<4>;
    else
        return (int)cudaErrorInvalidValue;  // k > 512: not supported by the ring kernel
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, 0, (cudaStream_t)stream);
}

int mfk_sgd_exact_level(const mfk_node *R, const unsigned *order, int count, float *P, float *Q, float *PG,
                        float *QG, int k_al, float lambda_p, float lambda_q, float eta, int slow_only,
                        float *e2_out, void *stream) {
    if (count <= 0) return 0;
    int rc = ensure_table();
    if (rc) return rc;
    k_sgd_exact_level<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        R, order, count, P, Q, PG, QG, k_al, lambda_p, lambda_q, eta, slow_only, e2_out);
    return (int)cudaGetLastError();
}

int mfk_sum_f32(const float *x, long long n, double *out1, void *stream) {
    k_sum_f32<<<grid_for(n, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(x, n, out1);
    return (int)cudaGetLastError();
}

int mfk_reg2(const float *M, const int *omega, int rows, int k_al, double *out1, void *stream) {
    if (rows <= 0) return 0;
    k_reg2<<<grid_for(rows, 128, 148 * 8), 128, 0, (cudaStream_t)stream>>>(M, omega, rows, k_al, out1);
    return (int)cudaGetLastError();
}

int mfk_finalize_rows(const float *M, const int *map, int rows, int k, int k_al, float factor, float *out,
                      void *stream) {
    if ((long long)rows * k == 0) return 0;
    k_finalize_rows<<<grid_for((long long)rows * k, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        M, map, rows, k, k_al, factor, out);
    return (int)cudaGetLastError();
}

int mfk_predict_pairs(const float *P, const float *Q, int m, int n, int k, float b, const float *pairs,
                      long long npairs, float *out, void *stream) {
    if (npairs <= 0) return 0;
    k_predict_pairs<<<grid_for(npairs, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(P, Q, m, n, k, b, pairs,
                                                                                      npairs, out);
    return (int)cudaGetLastError();
}

int mfk_sq_err(const mfk_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k, float b,
               double *out1, void *stream) {
    if (nnz <= 0) return 0;
    k_sq_err<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(R, nnz, P, Q, m, n, k, b, out1);
    return (int)cudaGetLastError();
}

}  // extern "C"
