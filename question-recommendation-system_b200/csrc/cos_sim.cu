// csrc/cos_sim.cu -- cosine similarity of the rows of an integer Q matrix (items x knowledge points), sm_100a.
//
// Replaces the arithmetic of mf::cos_similarity (mf/mf.cpp:3591-3683): for item a and every item i
//     cos[i] = (int) <q_a, q_i>  /  ( sqrt((double) <q_a, q_a>) * sqrt((double) <q_i, q_i>) )     -> float
// with the dot products in int arithmetic (mf/mf.cpp:3634-3649), then the items ordered by falling cosine.  The
// reference answers ONE item per call with an O(items^2) exchange sort; here one launch computes the cosines of a whole
// batch of items against all items (all items at once), one warp per (a, i) pair row, and a
// segmented radix sort orders every row.  Values are bit-exact (IEEE double sqrt / divide, one rounding to float).
// Compiled without -ftz like the other evaluation kernels.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>

#include "kernels.h"

namespace {

// one thread per (a, i): k is small (knowledge points), Q rows are read through L1/L2
__global__ void __launch_bounds__(256)
k_cos_rows(const int *__restrict__ Q, int items, int k, const int *__restrict__ a_list, int a_count, float *cos_out,
           int *id_out) {
    const long long total = (long long)a_count * items;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ai = (int)(t / items), i = (int)(t - (long long)ai * items);
        const int a = a_list ? a_list[ai] : ai;
        const int *qa = Q + (size_t)a * k, *qi = Q + (size_t)i * k;
        int item_abs = 0, every_abs = 0, dot = 0;  // int arithmetic, wraps like the reference's
        for (int d = 0; d < k; d++) {
            const int x = qa[d], y = qi[d];
            item_abs += x * x;
            dot += x * y;
            every_abs += y * y;
        }
        // mf/mf.cpp:3647: int / (double * double), stored to a float
        const float c = (float)((double)dot / (sqrt((double)item_abs) * sqrt((double)every_abs)));
        // a zero row gives 0/0: stored as a NaN with the sign bit set, which the descending radix sort puts last
        cos_out[t] = c != c ? __int_as_float((int)0xffc00000u) : c;
        id_out[t] = i;
    }
}

// after the sort: does a row hold NaNs or equal neighbours?  (then the reference's exchange sort decides the order,
// see mf_api.cpp; with distinct values every correct sort gives the same list)
__global__ void __launch_bounds__(256)
k_cos_flag_ties(const float *__restrict__ sorted, int items, int a_count, int *flags) {
    const long long total = (long long)a_count * items;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ai = (int)(t / items), i = (int)(t - (long long)ai * items);
        const float v = sorted[t];
        if (v != v || (i > 0 && sorted[t - 1] == v)) flags[ai] = 1;
    }
}

__global__ void k_cos_offsets(long long *offs, int items, int a_count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= a_count) offs[r] = (long long)r * items;
}

}  // namespace

extern "C" {

size_t mfk_cos_tmp_bytes(int items, int a_count) {
    size_t bytes = 0;
    const long long *off = nullptr;
    cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, bytes, (const float *)nullptr, (float *)nullptr,
                                                      (const int *)nullptr, (int *)nullptr, (long long)a_count * items, a_count,
                                                      off, off + 1);
    // + the row offsets of the segmented sort, kept in front of cub's own scratch space
    return ((bytes + 255) & ~(size_t)255) + (((size_t)(a_count + 1) * sizeof(long long) + 255) & ~(size_t)255);
}

// Q: dense [items][k] int matrix on the device.  a_list (device, may be NULL = items 0..a_count-1).  Outputs on the
// device: cos_raw [a_count][items] in item order, cos_sorted / id_sorted [a_count][items] by falling cosine (equal
// cosines: rising id), tie_flags [a_count] (must be zeroed by the caller).
int mfk_cos_similarity(const int *Q, int items, int k, const int *a_list, int a_count, float *cos_raw, int *id_raw,
                       float *cos_sorted, int *id_sorted, int *tie_flags, void *tmp, size_t tmp_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)a_count * items;
    if (total <= 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    k_cos_rows<<<grid, 256, 0, st>>>(Q, items, k, a_list, a_count, cos_raw, id_raw);
    long long *off = (long long *)tmp;
    const size_t off_bytes = ((size_t)(a_count + 1) * sizeof(long long) + 255) & ~(size_t)255;
    if (tmp_bytes <= off_bytes) return (int)cudaErrorInvalidValue;
    k_cos_offsets<<<(a_count + 256) / 256, 256, 0, st>>>(off, items, a_count);
    size_t cub_bytes = tmp_bytes - off_bytes;
    cudaError_t e = cub::DeviceSegmentedRadixSort::SortPairsDescending((char *)tmp + off_bytes, cub_bytes, cos_raw, cos_sorted,
                                                                      id_raw, id_sorted, total, a_count, off, off + 1, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    k_cos_flag_ties<<<grid, 256, 0, st>>>(cos_sorted, items, a_count, tie_flags);
    return (int)cudaGetLastError();
}

}  // extern "C"
