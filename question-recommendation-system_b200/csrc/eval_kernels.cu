// csrc/eval_kernels.cu -- the predict path and the error measures, sm_100a.
//
//   k_predict_pairs   mf::utility_predict -> mf_predict (mf/mf.cpp:3537-3568, 4295-4314)
//   k_sq_err          calc_rmse's sum (mf/mf.cpp:4316-4331)
//   k_va_err          the validation column of fpsg_core's table (mf/mf.cpp:2884-2904, calc_error 635-660)
//   k_err_general     calc_mae / calc_gkl / calc_logloss / calc_accuracy (4333-4404), calc_error (635-674), the
//                     cross-validation error over hidden blocks (2918-2938)
//
// Every value is bit-exact to the reference: z is the sequential fp32 sum over the dimensions in index order, the
// product rounded before the add (the reference is built without FMA, mf/CMakeLists.txt:10).  This file is compiled
// WITHOUT -ftz: mf_predict and the calc_* functions run outside fpsg_core's flush-to-zero window
// (mf/mf.cpp:2789-2790, 2941), so denormal products and sums are kept like on the CPU.
//
// Memory access (round 1 let one thread walk its own two rows 4 bytes at a time -- 32 lanes on 32 different rows):
// a warp takes 32 pairs; for every block of 32 dimensions it stages the 32 + 32 row pieces in shared memory with
// 128-bit loads, four full 128-byte segments per instruction, and every lane then sums ITS OWN pair sequentially from
// shared memory (128-bit reads, conflict-free with a row pitch of 36 floats) -- the order of the additions is untouched.
// Roofline: HBM, 2 * 4k bytes per pair (both rows), reported by bench.py.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kDimBlock = 32;              // dimensions staged at a time
constexpr int kPitch = kDimBlock + 4;      // floats per staged row piece: 16-byte aligned, conflict-free LDS.128
constexpr int kEvalWarps = 4;              // warps per CTA
constexpr int kTileFloats = 2 * 32 * kPitch;  // per warp: 32 row pieces of P and of Q

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

// mf_predict for the 32 pairs of a warp (lane i owns pair (u, v); `in` = both ids in range).  All lanes must call it.
// stride = floats per row (k for a finished model, k_al for the training-space model), k = dimensions summed.
// Rows must be 16-byte aligned (stride % 4 == 0); the caller falls back to predict_scalar otherwise.
__device__ __forceinline__ float predict_warp32(const float *__restrict__ P, const float *__restrict__ Q, int stride, int k,
                                                float b, int u, int v, bool in, float *tile /* kTileFloats */) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7, quad = lane >> 3;  // 8 lanes x 16 bytes = one 128-byte row piece; four pairs per instruction
    float *tp = tile, *tq = tile + 32 * kPitch;
    // row pointers of the four pairs this lane helps to load in each of the eight rounds
    const float *pj[8], *qj[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = r * 4 + quad;
        const int uj = __shfl_sync(kFull, u, j), vj = __shfl_sync(kFull, v, j);
        const bool inj = __shfl_sync(kFull, (int)in, j) != 0;
        pj[r] = inj ? P + (size_t)uj * stride : nullptr;
        qj[r] = inj ? Q + (size_t)vj * stride : nullptr;
    }
    float z = 0.0f;
    for (int d0 = 0; d0 < k; d0 += kDimBlock) {
        const int nd = min(kDimBlock, k - d0);
        __syncwarp();  // the previous block has been consumed
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int j = r * 4 + quad;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
            if (pj[r] && sub * 4 < nd) {  // (k % 4 == 0 here, so a 16-byte piece is entirely inside or outside the row)
                a = __ldg(reinterpret_cast<const float4 *>(pj[r] + d0) + sub);
                c = __ldg(reinterpret_cast<const float4 *>(qj[r] + d0) + sub);
            }
            *reinterpret_cast<float4 *>(tp + j * kPitch + sub * 4) = a;
            *reinterpret_cast<float4 *>(tq + j * kPitch + sub * 4) = c;
        }
        __syncwarp();
        const float4 *mp = reinterpret_cast<const float4 *>(tp + lane * kPitch), *mq = reinterpret_cast<const float4 *>(tq + lane * kPitch);
#pragma unroll
        for (int i = 0; i < kDimBlock / 4; i++) {
            if (i * 4 < nd) {
                const float4 a = mp[i], c = mq[i];
                z = __fadd_rn(z, __fmul_rn(a.x, c.x));
                z = __fadd_rn(z, __fmul_rn(a.y, c.y));
                z = __fadd_rn(z, __fmul_rn(a.z, c.z));
                z = __fadd_rn(z, __fmul_rn(a.w, c.w));
            }
        }
    }
    return (!in || isnan(z)) ? b : z;  // out of range -> b; NaN (a row never trained on) -> b: mf/mf.cpp:4297-4312
}

// the same value with plain scalar loads, for strides that are not a multiple of four floats
__device__ __forceinline__ float predict_scalar(const float *__restrict__ P, const float *__restrict__ Q, int stride, int k,
                                                float b, int u, int v, bool in) {
    if (!in) return b;
    const float *p = P + (size_t)u * stride, *q = Q + (size_t)v * stride;
    float z = 0.0f;
    for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(p[d], q[d]));
    return isnan(z) ? b : z;
}

// VEC: rows are 16-byte aligned.  The loop bound is warp-uniform (every lane runs every round; a lane past the end
// carries a pair that is "not in range" and throws its result away), so predict_warp32's shuffles are legal.
template <bool VEC>
__device__ __forceinline__ float predict_any(const float *P, const float *Q, int stride, int k, float b, int u, int v,
                                             bool in, float *tile) {
    return VEC ? predict_warp32(P, Q, stride, k, b, u, v, in, tile) : predict_scalar(P, Q, stride, k, b, u, v, in);
}
__device__ __forceinline__ long long rounds_for(long long n) {  // rounds of gridDim * blockDim items
    const long long per = (long long)gridDim.x * blockDim.x;
    return (n + per - 1) / per;
}

template <bool VEC>
__global__ void __launch_bounds__(kEvalWarps * 32)
k_predict_pairs(const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k, float b,
                const float *__restrict__ pairs, long long npairs, float *out) {
    __shared__ __align__(16) float tiles[kEvalWarps][kTileFloats];
    float *tile = tiles[threadIdx.x >> 5];
    const long long per = (long long)gridDim.x * blockDim.x, nr = rounds_for(npairs);
    for (long long rd = 0; rd < nr; rd++) {
        const long long i = rd * per + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        int u = -1, v = -1;
        if (i < npairs) {  // ids travel as floats (mf/mf.cpp:3562-3565)
            const float2 uv = __ldg(reinterpret_cast<const float2 *>(pairs) + i);
            u = (int)uv.x;
            v = (int)uv.y;
        }
        const bool in = u >= 0 && u < m && v >= 0 && v < n;
        const float z = predict_any<VEC>(P, Q, k, k, b, u, v, in, tile);
        if (i < npairs) out[i] = z;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(kEvalWarps * 32)
k_sq_err(const mfk_node *__restrict__ R, long long nnz, const float *__restrict__ P, const float *__restrict__ Q, int m,
         int n, int k, float b, double *out) {
    __shared__ __align__(16) float tiles[kEvalWarps][kTileFloats];
    __shared__ double sm[32];
    float *tile = tiles[threadIdx.x >> 5];
    double s = 0.0;
    const long long per = (long long)gridDim.x * blockDim.x, nr = rounds_for(nnz);
    for (long long rd = 0; rd < nr; rd++) {
        const long long i = rd * per + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        mfk_node N = {-1, -1, 0.f};
        if (i < nnz) N = R[i];
        const bool in = N.u >= 0 && N.u < m && N.v >= 0 && N.v < n;
        const float z = predict_any<VEC>(P, Q, k, k, b, N.u, N.v, in, tile);
        if (i < nnz) {
            const float e = __fsub_rn(N.r, z);
            s += (double)__fmul_rn(e, e);  // calc_rmse: (float)(e * e) widened, mf/mf.cpp:4325-4327
        }
    }
    s = block_sum_double(s, sm);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// The validation column of fpsg_core's table (mf/mf.cpp:2884-2904): calc_error (635-660) over the validation set in
// TRAINING space -- ids through the same permutations (shuffle_problem, 775-791: ids beyond the map are kept),
// r * 1/scale (scale_problem), z = mf_predict on the k_al-strided model, error += pow(r - z, 2) in double.
__global__ void __launch_bounds__(kEvalWarps * 32)
k_va_err(const mfk_node *__restrict__ R, long long nnz, const int *__restrict__ p_map, const int *__restrict__ q_map,
         const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k_al, float b, float inv_scale,
         double *out) {
    __shared__ __align__(16) float tiles[kEvalWarps][kTileFloats];
    __shared__ double sm[32];
    float *tile = tiles[threadIdx.x >> 5];
    double s = 0.0;
    const long long per = (long long)gridDim.x * blockDim.x, nr = rounds_for(nnz);
    for (long long rd = 0; rd < nr; rd++) {
        const long long i = rd * per + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        mfk_node N = {-1, -1, 0.f};
        if (i < nnz) N = R[i];
        const int u = (N.u >= 0 && N.u < m) ? p_map[N.u] : N.u, v = (N.v >= 0 && N.v < n) ? q_map[N.v] : N.v;
        const bool in = u >= 0 && u < m && v >= 0 && v < n;
        const float z = predict_warp32(P, Q, k_al, k_al, b, u, v, in, tile);  // k_al is a multiple of 8
        if (i < nnz) {
            const float r = inv_scale == 1.0f ? N.r : __fmul_rn(N.r, inv_scale);
            const double d = (double)__fsub_rn(r, z);
            s += d * d;
        }
    }
    s = block_sum_double(s, sm);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// The other error measures: calc_mae / calc_gkl / calc_logloss / calc_accuracy (mf/mf.cpp:4333-4404) on a finished
// model, and calc_error (635-674) on the training-space model for the validation column (p_map != NULL: ids through
// the permutations, r * 1/scale).  `which` uses the loss codes: 1 sum |r - z|, 2 sum r log(r/z) - r + z, 5 sum
// log(1 + exp(-+z)) in double, 6/7 number of correctly classified ratings, otherwise sum (r - z)^2.
template <bool VEC>
__global__ void __launch_bounds__(kEvalWarps * 32)
k_err_general(int which, const mfk_node *__restrict__ R, long long nnz, const int *__restrict__ p_map,
              const int *__restrict__ q_map, const float *__restrict__ P, const float *__restrict__ Q, int m, int n,
              int k, float b, float inv_scale, double *out, int train_space, mfk_hidden hid) {
    __shared__ __align__(16) float tiles[kEvalWarps][kTileFloats];
    __shared__ double sm[32];
    float *tile = tiles[threadIdx.x >> 5];
    double s = 0.0;
    unsigned long long used = 0;
    const long long per = (long long)gridDim.x * blockDim.x, nr = rounds_for(nnz);
    for (long long rd = 0; rd < nr; rd++) {
        const long long i = rd * per + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        mfk_node N = {-1, -1, 0.f};
        if (i < nnz) N = R[i];
        int u = N.u, v = N.v;
        float r = N.r;
        if (p_map) {
            u = (u >= 0 && u < m) ? p_map[u] : u;
            v = (v >= 0 && v < n) ? q_map[v] : v;
            if (inv_scale != 1.0f) r = __fmul_rn(r, inv_scale);
        }
        const bool in = u >= 0 && u < m && v >= 0 && v < n;
        // cross-validation error: only the ratings of the hidden grid blocks (ids in training space here)
        const bool take = i < nnz && !(hid.mask && !(in && hid.mask[(u / hid.seg_p) * hid.bins + v / hid.seg_q]));
        const float z = predict_any<VEC>(P, Q, k, k, b, u, v, in, tile);
        if (!take) continue;  // (after the warp-wide call)
        used++;
        switch (which) {
            case MFK_FUN_L1_MFR: s += (double)fabsf(__fsub_rn(r, z)); break;
            case MFK_FUN_KL_MFR:
                s += (double)__fadd_rn(__fsub_rn(__fmul_rn(r, (float)log((double)__fdiv_rn(r, z))), r), z);
                break;
            case MFK_FUN_LR_MFC:
                s += r > 0.f ? log(1.0 + (double)(float)exp((double)-z)) : log(1.0 + (double)(float)exp((double)z));
                break;
            case MFK_FUN_L2_MFC:
            case MFK_FUN_L1_MFC: s += r > 0.f ? (z > 0.f ? 1.0 : 0.0) : (z < 0.f ? 1.0 : 0.0); break;
            default: {
                if (train_space) {  // calc_error: pow(r - z, 2) on the double
                    const double d = (double)__fsub_rn(r, z);
                    s += d * d;
                } else {  // calc_rmse: (float)(e * e)
                    const float e = __fsub_rn(r, z);
                    s += (double)__fmul_rn(e, e);
                }
            }
        }
    }
    s = block_sum_double(s, sm);
    if (threadIdx.x == 0) atomicAdd(out, s);
    if (hid.mask) {  // out[1] += number of ratings that took part
        __syncthreads();
        const double c = block_sum_double((double)used, sm);
        if (threadIdx.x == 0) atomicAdd(out + 1, c);
    }
}

inline int eval_grid(long long n) {
    long long g = (n + kEvalWarps * 32 - 1) / (kEvalWarps * 32);
    if (g < 1) g = 1;
    if (g > 148 * 24) g = 148 * 24;
    return (int)g;
}
inline bool rows_aligned(const float *P, const float *Q, int stride) {
    return stride % 4 == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0;
}

}  // namespace

extern "C" {

int mfk_err_general(int which, const mfk_node *R, long long nnz, const int *p_map, const int *q_map, const float *P,
                    const float *Q, int m, int n, int k, float b, float inv_scale, double *out1, int train_space,
                    mfk_hidden hidden, void *stream) {
    if (nnz <= 0) return 0;
    if (rows_aligned(P, Q, k))
        k_err_general<true><<<eval_grid(nnz), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(
            which, R, nnz, p_map, q_map, P, Q, m, n, k, b, inv_scale, out1, train_space, hidden);
    else
        k_err_general<false><<<eval_grid(nnz), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(
            which, R, nnz, p_map, q_map, P, Q, m, n, k, b, inv_scale, out1, train_space, hidden);
    return (int)cudaGetLastError();
}

int mfk_predict_pairs(const float *P, const float *Q, int m, int n, int k, float b, const float *pairs,
                      long long npairs, float *out, void *stream) {
    if (npairs <= 0) return 0;
    if (rows_aligned(P, Q, k))
        k_predict_pairs<true><<<eval_grid(npairs), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(P, Q, m, n, k, b, pairs,
                                                                                              npairs, out);
    else
        k_predict_pairs<false><<<eval_grid(npairs), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(P, Q, m, n, k, b, pairs,
                                                                                               npairs, out);
    return (int)cudaGetLastError();
}

int mfk_va_err(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, const float *P, const float *Q,
               int m, int n, int k_al, float b, float inv_scale, double *out1, void *stream) {
    if (nnz <= 0) return 0;
    k_va_err<<<eval_grid(nnz), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(R, nnz, p_map, q_map, P, Q, m, n, k_al, b,
                                                                          inv_scale, out1);
    return (int)cudaGetLastError();
}

int mfk_sq_err(const mfk_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k, float b,
               double *out1, void *stream) {
    if (nnz <= 0) return 0;
    if (rows_aligned(P, Q, k))
        k_sq_err<true><<<eval_grid(nnz), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(R, nnz, P, Q, m, n, k, b, out1);
    else
        k_sq_err<false><<<eval_grid(nnz), kEvalWarps * 32, 0, (cudaStream_t)stream>>>(R, nnz, P, Q, m, n, k, b, out1);
    return (int)cudaGetLastError();
}

}  // extern "C"
