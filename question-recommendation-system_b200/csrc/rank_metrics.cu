// csrc/rank_metrics.cu -- the ranking measures of the one-class losses, sm_100a: calc_mpr_auc (mf/mf.cpp:4406-4525).
//
// For every row i (a user; an item when `transpose`) that has positives: the score of EVERY column by mf_predict
// (4295-4314: sequential fp32 sum from 0.0f, product rounded before the add, NaN -> b, out of range -> b); the positives'
// scores sorted ascending; for every other column the number `left` of positives that do not beat it (the binary search of
// 4480-4498 = an upper bound); u_mpr = sum left, u_auc = sum (pos - left).  The reference then adds
// u_mpr / (n - pos) and u_auc / (n - pos) / pos over the rows in double.  Here the two sums of a row are exact integers
// computed on the device, and the host adds the rows' terms in rising row order -- the reference's order at one thread -- so
// the two measures equal the reference's to the last bit.  Compiled without -ftz like the other evaluation kernels.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>

#include "kernels.h"

namespace {

// key = row << 32 | column for ratings with r > 0, ~0 otherwise (they sort to the end and are not counted)
__global__ void __launch_bounds__(256)
k_rank_keys(const mfk_node *__restrict__ R, long long nnz, int transpose, unsigned long long *keys, unsigned long long *kept) {
    unsigned long long mine = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
        const mfk_node N = R[i];
        const unsigned row = (unsigned)(transpose ? N.v : N.u), col = (unsigned)(transpose ? N.u : N.v);
        const bool ok = N.r > 0.f && N.u >= 0 && N.v >= 0;
        keys[i] = ok ? ((unsigned long long)row << 32) | col : ~0ull;
        mine += ok ? 1ull : 0ull;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept, mine);
}

// row_start[r] = first sorted position of row r (rows without positives get the next row's start)
__global__ void __launch_bounds__(256)
k_rank_row_starts(const unsigned long long *__restrict__ keys, long long npos, int rows, long long *row_start) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npos; i += (long long)gridDim.x * blockDim.x) {
        const long long row = (long long)(keys[i] >> 32), prev = i > 0 ? (long long)(keys[i - 1] >> 32) : -1;
        for (long long r = prev + 1; r <= row && r <= rows; r++) row_start[r] = i;
        if (i == npos - 1)
            for (long long r = row + 1; r <= rows; r++) row_start[r] = npos;
    }
}

// scores[(row - lo) * cols + col] = mf_predict(row, col)
__global__ void __launch_bounds__(256)
k_rank_scores(const float *__restrict__ P, const float *__restrict__ Q, int m, int n, int k, float b, int transpose, int lo,
              int hi, int cols, float *scores) {
    const long long total = (long long)(hi - lo) * cols;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int row = lo + (int)(t / cols), col = (int)(t - (long long)(row - lo) * cols);
        const int u = transpose ? col : row, v = transpose ? row : col;
        float z = b;
        if (u < m && v < n) {
            const float *p = P + (size_t)u * k, *q = Q + (size_t)v * k;
            z = 0.0f;
            for (int d = 0; d < k; d++) z = __fadd_rn(z, __fmul_rn(p[d], q[d]));
            if (z != z) z = b;
        }
        scores[t] = z;
    }
}

// the positives of the batch: their scores go to pos_scores (sorted positions), their place in `scores` is marked NaN
__global__ void __launch_bounds__(256)
k_rank_take_positives(const unsigned long long *__restrict__ keys, long long first, long long last, int lo, int cols,
                      float *scores, float *pos_scores) {
    for (long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < last; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(keys[i] >> 32), col = (int)(unsigned)keys[i];
        float *s = scores + (size_t)(row - lo) * cols + col;
        pos_scores[i] = *s;
        *s = __int_as_float(0x7fc00000);
    }
}

// one CTA per row: for every column that is not a positive, left = number of positives with score <= its score
__global__ void __launch_bounds__(256)
k_rank_count(const float *__restrict__ scores, const float *__restrict__ pos_sorted, const long long *__restrict__ row_start,
             int lo, int cols, unsigned long long *u_mpr, unsigned long long *u_auc) {
    const int row = lo + blockIdx.x;
    const long long s0 = row_start[row], s1 = row_start[row + 1];
    const int pos = (int)(s1 - s0);
    unsigned long long a = 0, c = 0;
    if (pos > 0) {
        const float *ps = pos_sorted + s0;
        const float *sc = scores + (size_t)blockIdx.x * cols;
        for (int j = threadIdx.x; j < cols; j += blockDim.x) {
            const float s = sc[j];
            if (s != s) continue;  // a positive
            int l = 0, r = pos;    // first index with ps[idx] > s
            while (l < r) {
                const int mid = (l + r) >> 1;
                if (ps[mid] > s) r = mid; else l = mid + 1;
            }
            a += (unsigned long long)l;
            c += (unsigned long long)(pos - l);
        }
    }
    __shared__ unsigned long long sa[8], sc2[8];
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sa[threadIdx.x >> 5] = a;
        sc2[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ta = 0, tc = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            ta += sa[w];
            tc += sc2[w];
        }
        u_mpr[row] = ta;
        u_auc[row] = tc;
    }
}

}  // namespace

extern "C" {

size_t mfk_rank_sort_tmp_bytes(long long nnz) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, nnz);
    const long long *off = nullptr;
    cub::DeviceSegmentedRadixSort::SortKeys(nullptr, b, (const float *)nullptr, (float *)nullptr, nnz, 1, off, off + 1);
    return (a > b ? a : b) + 256;
}

// step 1: keys of the positives, sorted by (row, column); *kept_dev = their number; row_start[rows + 1]
int mfk_rank_prepare(const mfk_node *R, long long nnz, int transpose, int rows, unsigned long long *keys_tmp,
                     unsigned long long *keys_sorted, unsigned long long *kept_dev, long long *row_start, void *tmp,
                     size_t tmp_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(kept_dev, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    const int grid = (int)std::min<long long>((nnz + 255) / 256, 148 * 8);
    k_rank_keys<<<grid, 256, 0, st>>>(R, nnz, transpose, keys_tmp, kept_dev);
    e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_tmp, keys_sorted, nnz, 0, 64, st);
    if (e != cudaSuccess) return (int)e;
    unsigned long long kept = 0;
    e = cudaMemcpyAsync(&kept, kept_dev, sizeof(kept), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(row_start, 0, sizeof(long long) * ((size_t)rows + 1), st);  // no positives at all: all zero
    if (e != cudaSuccess) return (int)e;
    if (kept > 0) k_rank_row_starts<<<grid, 256, 0, st>>>(keys_sorted, (long long)kept, rows, row_start);
    return (int)cudaGetLastError();
}

// step 2, for rows [lo, hi): scores, positives' scores sorted per row, the two integer sums of every row.
// first / last: sorted positions of the positives of these rows (row_start[lo], row_start[hi], host copies).
int mfk_rank_batch(const float *P, const float *Q, int m, int n, int k, float b, int transpose, int lo, int hi, int cols,
                   const unsigned long long *keys_sorted, const long long *row_start, long long first, long long last,
                   float *scores, float *pos_scores, float *pos_sorted, unsigned long long *u_mpr, unsigned long long *u_auc,
                   void *tmp, size_t tmp_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (hi <= lo) return 0;
    const long long total = (long long)(hi - lo) * cols;
    k_rank_scores<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(P, Q, m, n, k, b, transpose, lo, hi,
                                                                                         cols, scores);
    if (last > first) {
        k_rank_take_positives<<<(int)std::min<long long>((last - first + 255) / 256, 148 * 8), 256, 0, st>>>(
            keys_sorted, first, last, lo, cols, scores, pos_scores);
        cudaError_t e = cub::DeviceSegmentedRadixSort::SortKeys(tmp, tmp_bytes, pos_scores, pos_sorted, last, hi - lo,
                                                               row_start + lo, row_start + lo + 1, 0, 32, st);
        if (e != cudaSuccess) return (int)e;
    }
    k_rank_count<<<hi - lo, 256, 0, st>>>(scores, pos_sorted, row_start, lo, cols, u_mpr, u_auc);
    return (int)cudaGetLastError();
}

}  // extern "C"
