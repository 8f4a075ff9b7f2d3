// csrc/kernels.cu -- sm_100a kernels of the matrix-factorisation hot path + their launchers.
//
// Kernels (reference lines they replace are cited at each one; paths relative to /root/reference):
//   k_sgd_band_epoch   throughput SGD: 8 lanes per rating update, float4 row accesses, shuffle dot
//                      product fused with the regularised AdaGrad step; the smaller factor matrix is
//                      resident in shared memory band by band, the other one streams; conflict-free
//                      schedule (band ring over CTAs, per-row tickets inside a CTA)
//   k_sgd_exact_level  bit-exact SGD in the reference's sequential order, one wavefront per launch
//   k_stats / k_band_keys2 / k_band_heads / k_band_keys1 / k_band_stream / k_init_rows / k_finalize_rows
//                      preprocessing of fpsg
//   k_reg1 / k_reg2                                                        the objective column
//   (predict + metrics: eval_kernels.cu, compiled without flush-to-zero)
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -ftz=true (the reference runs its loop with
// flush-to-zero on, mf/mf.cpp:2788-2791).  Exact kernels use __f*_rn intrinsics so that nothing
// is contracted into an FMA (the reference is built without FMA, mf/CMakeLists.txt:10).

#include <algorithm>
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "rsqrt12_table.h"
#include "dev_helpers.cuh"

namespace {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

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
// band preprocessing (replaces shuffle_problem / scale_problem / grid_problem, mf/mf.cpp:775-791,
// 517-527, 793-858, for the throughput schedule).  Coordinates of a rating (a = T row, b = S row):
//   js  stripe of b          sb  S band inside the stripe     bl  row inside the band
//   tb  T band of a          ai  row inside the T band        ga  group (T sub-band) = ai mod nG
//   c   CTA = sb mod nC      t   step at which CTA c meets T band tb = (tb - c*S1) mod nTB
//   d   phase = (bl mod nG - ga) mod nG: the order in which a group walks the items of its band, so
//       that at any moment different groups of a CTA tend to work on different items
// ------------------------------------------------------------------------------------------------
// item kernel: where the walk over the T rows starts for S row b
__device__ __forceinline__ unsigned mfk_item_rot(unsigned b, unsigned t_rows) {
    return (unsigned)(((unsigned long long)b * 2654435761ull) % (unsigned long long)t_rows);
}
struct BandCoord {
    unsigned js, sb, bl, tb, ai, ga, t, d;
};
__device__ __forceinline__ BandCoord band_coord(const mfk_band_shape &sh, unsigned a_local, unsigned b) {
    BandCoord x;
    x.js = b / (unsigned)sh.stripeRows;
    const unsigned bs = b - x.js * (unsigned)sh.stripeRows;
    x.sb = bs / (unsigned)sh.segS;
    x.bl = bs - x.sb * (unsigned)sh.segS;
    x.tb = a_local / (unsigned)sh.segT;
    x.ai = a_local - x.tb * (unsigned)sh.segT;
    // rows of a T band are dealt to the groups round-robin: no group stays empty (cells, by_row == 2: no group field)
    x.ga = sh.by_row == 2 ? 0u : x.ai % (unsigned)sh.nG;
    const unsigned c = x.sb % (unsigned)sh.nC;
    x.t = (x.tb + (unsigned)sh.nTB - (c * (unsigned)sh.S1) % (unsigned)sh.nTB) % (unsigned)sh.nTB;
    x.d = sh.by_row ? 0u : (x.bl % (unsigned)sh.nG + (unsigned)sh.nG - x.ga) % (unsigned)sh.nG;
    if (sh.by_row == 4) {
        // item kernel: the group is the owner of the S ROW (bl mod nG), the "step" orders that group's S rows, and the
        // walk over the T rows of an S row starts at a place that depends on the row (mfk_item_rot)
        x.ga = x.bl % (unsigned)sh.nG;
        x.t = x.bl / (unsigned)sh.nG;
        x.ai = (a_local + mfk_item_rot(b, (unsigned)sh.tRows)) % (unsigned)sh.tRows;
    }
    return x;
}

// stream order: key1 = (stripe, S band) | group | step | phase | row inside the T band; payload = b | r/scale
__global__ void __launch_bounds__(256)
k_band_keys1(const mfk_node *__restrict__ R, long long nnz, const int *__restrict__ p_map,
             const int *__restrict__ q_map, mfk_band_shape sh, float inv_scale, int *omega_p, int *omega_q,
             unsigned long long *keys, unsigned long long *vals, unsigned long long *kept, int *bad, int m, int n,
             mfk_hidden hid) {
    unsigned long long mine = 0;
    const unsigned nBands = (unsigned)sh.nC * (unsigned)sh.nPass;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
         i += (long long)gridDim.x * blockDim.x) {
        const mfk_node N = R[i];
        unsigned long long key = ~0ull, val = 0ull;
        if (N.u < 0 || N.u >= m || N.v < 0 || N.v >= n) {
            *bad = 1;
        } else {
            const int u = p_map[N.u], v = q_map[N.v];
            const int a = sh.swap_sides ? v : u, b = sh.swap_sides ? u : v;
            if (omega_p) {  // omega counts ALL ratings (NULL: already counted by k_owner_of, sharded load)
                atomicAdd(omega_p + u, 1);
                atomicAdd(omega_q + v, 1);
            }
            // cross-validation: ratings of hidden grid blocks are counted (omega) but never trained on
            const bool hidden = hid.mask && hid.mask[(u / hid.seg_p) * hid.bins + v / hid.seg_q];
            if (!hidden && a >= sh.tLo && a < sh.tLo + sh.tRows) {
                const BandCoord x = band_coord(sh, (unsigned)(a - sh.tLo), (unsigned)b);
                key = ((unsigned long long)(x.js * nBands + x.sb) << sh.bitsG) | x.ga;
                key = (((key << sh.bitsT | x.t) << sh.bitsD | x.d) << sh.bitsA) | x.ai;
                // scale_problem (mf/mf.cpp:517-527): r * (1/scale), skipped when the factor is exactly 1
                const float r = inv_scale == 1.0f ? N.r : __fmul_rn(N.r, inv_scale);
                val = ((unsigned long long)(unsigned)b << 32) | __float_as_uint(r);
                mine++;
            }
        }
        keys[i] = key;
        vals[i] = val;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFull, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept, mine);
}

// Sharded load (several GPUs): every rank reads only its slice of the caller's rating array and sends each rating to
// the rank that owns its T row.  owner[i] = T row / rows per rank (the sort key that groups the slice by destination),
// counts[owner]++; omega is counted here, on the slice, and summed over the ranks afterwards.
__global__ void __launch_bounds__(256)
k_owner_of(const mfk_node *__restrict__ R, long long nnz, const int *__restrict__ p_map, const int *__restrict__ q_map,
           int swap_sides, int t_seg, int world, int *omega_p, int *omega_q, unsigned char *owner,
           unsigned long long *counts, int *bad, int m, int n) {
    // counts[] has only `world` words: one global atomic per rating on them serialises (measured: 22 ms for 50M ratings
    // and two owners).  Lanes with the same owner are counted once per warp into a per-block histogram, which reaches
    // global memory once per block and owner.
    __shared__ unsigned s_cnt[256];
    s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    // block-uniform trip count, so that the warp votes below see full warps
    for (long long base = blockIdx.x * (long long)blockDim.x; base < nnz; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        unsigned o = 255u;  // no rating in this lane
        if (i < nnz) {
            const mfk_node N = R[i];
            o = (unsigned)world;  // out of range: goes nowhere
            if (N.u < 0 || N.u >= m || N.v < 0 || N.v >= n) {
                *bad = 1;
            } else {
                const int u = p_map[N.u], v = q_map[N.v];
                atomicAdd(omega_p + u, 1);
                atomicAdd(omega_q + v, 1);
                o = (unsigned)min((swap_sides ? v : u) / t_seg, world - 1);
            }
            owner[i] = (unsigned char)o;
        }
        const unsigned peers = __match_any_sync(kFull, o);
        if (o < (unsigned)world && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_cnt[o], (unsigned)__popc(peers));
    }
    __syncthreads();
    if ((int)threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// head[i] = i where a new (S band, group, step) segment of the stream starts, else 0; an inclusive max-scan
// turns that into the segment start of every entry.
__global__ void __launch_bounds__(256)
k_band_seghead(const unsigned long long *__restrict__ k1, long long cnt, mfk_band_shape sh, unsigned *head) {
    const int sh_bits = sh.bitsD + sh.bitsA;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt;
         i += (long long)gridDim.x * blockDim.x)
        head[i] = (i > 0 && (k1[i] >> sh_bits) != (k1[i - 1] >> sh_bits)) ? (unsigned)i : 0u;
}
struct MaxU32 {
    __device__ __forceinline__ unsigned operator()(unsigned a, unsigned b) const { return a > b ? a : b; }
};

// ticket order of an S row: key2 = b | step | rank | group, where rank is the position of the rating inside
// its (group, step) segment.  A group's r-th rating of a step therefore has priority r on the row it
// touches, whatever the phase: groups that advance at the same pace do not wait for one another.
__global__ void __launch_bounds__(256)
k_band_keys2(const unsigned long long *__restrict__ k1, const unsigned long long *__restrict__ v1,
             const unsigned *__restrict__ segstart, long long cnt, mfk_band_shape sh, int bitsR,
             unsigned long long *k2, unsigned *idx) {
    const unsigned rmax = (1u << bitsR) - 1u;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = k1[i];
        const unsigned t = (unsigned)(key >> (sh.bitsA + sh.bitsD)) & ((1u << sh.bitsT) - 1u);
        const unsigned ga = (unsigned)(key >> (sh.bitsA + sh.bitsD + sh.bitsT)) & ((1u << sh.bitsG) - 1u);
        const unsigned b = (unsigned)(v1[i] >> 32);
        const unsigned rank = min((unsigned)i - segstart[i], rmax);
        k2[i] = ((((unsigned long long)b << sh.bitsT | t) << bitsR | rank) << sh.bitsG) | ga;
        idx[i] = (unsigned)i;
    }
}

__global__ void __launch_bounds__(256)
k_band_heads(const unsigned long long *__restrict__ k2, long long cnt, int shift, unsigned *first) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned b = (unsigned)(k2[i] >> shift);
        if (i == 0 || (unsigned)(k2[i - 1] >> shift) != b) first[b] = (unsigned)i;
    }
}

__global__ void __launch_bounds__(256)
k_band_tickets(const unsigned long long *__restrict__ k2, const unsigned *__restrict__ idx, long long cnt, int shift,
               const unsigned *__restrict__ first, unsigned *ticket) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt;
         i += (long long)gridDim.x * blockDim.x)
        ticket[idx[i]] = ((unsigned)i - first[(unsigned)(k2[i] >> shift)]) & MFK_TICKET_MASK;
}

__global__ void __launch_bounds__(256)
k_band_stream(const unsigned long long *__restrict__ k1, const unsigned long long *__restrict__ v1,
              const unsigned *__restrict__ ticket, long long cnt, mfk_band_shape sh, unsigned *w0, unsigned *w1,
              float *rr, unsigned *goff) {
    const int lowbits = sh.bitsT + sh.bitsD + sh.bitsA;
    const long long nOff = (long long)sh.nStripes * sh.nC * sh.nPass * (sh.by_row == 2 ? sh.nTB : sh.nG);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = k1[i], val = v1[i];
        const unsigned ai = (unsigned)key & ((1u << sh.bitsA) - 1u);
        const unsigned t = (unsigned)(key >> (sh.bitsA + sh.bitsD)) & ((1u << sh.bitsT) - 1u);
        const unsigned b = (unsigned)(val >> 32);
        const unsigned bs = b % (unsigned)sh.stripeRows, bl = bs % (unsigned)sh.segS;
        const unsigned long long hi = key >> lowbits;
        long long slot, prev = -1;
        if (sh.by_row == 2) {
            // cells (k_sgd_cell_epoch): no group field; offsets per (S band, step); run heads and continuations marked
            const unsigned cta = (unsigned)(hi % (unsigned long long)((unsigned)sh.nC * (unsigned)sh.nPass)) % (unsigned)sh.nC;
            const unsigned tb = (t + cta * (unsigned)sh.S1) % (unsigned)sh.nTB;
            const bool head = i == 0 || k1[i - 1] != key, cont = i + 1 < cnt && k1[i + 1] == key;
            w0[i] = (head ? MFK_CELL_HEAD : 0u) | (cont ? MFK_CELL_CONT : 0u) | (tb * (unsigned)sh.segT + ai);
            w1[i] = (t << MFK_W1_BBITS) | bl;
            slot = (long long)hi * sh.nTB + t;
            if (i > 0) {
                const unsigned long long pk = k1[i - 1];
                prev = (long long)(pk >> lowbits) * sh.nTB + (long long)((unsigned)(pk >> (sh.bitsA + sh.bitsD)) & ((1u << sh.bitsT) - 1u));
            }
        } else {
            if (sh.by_row == 4) {
                // item kernel: w0 = the T row itself (the rotation of the sort key undone), w1 = the S row inside its band
                w0[i] = (ai + (unsigned)sh.tRows - mfk_item_rot(b, (unsigned)sh.tRows)) % (unsigned)sh.tRows;
                w1[i] = bl;
            } else if (sh.by_row && !ticket) {
                // the run kernel with locks: w0 = the T row itself (relative to this rank's band), w1 = step | S row, so
                // the kernel spends nothing on decoding (no ticket is needed: its field carries the step)
                const unsigned cta = (unsigned)((hi >> sh.bitsG) % (unsigned long long)((unsigned)sh.nC * (unsigned)sh.nPass)) % (unsigned)sh.nC;
                const unsigned tb = (t + cta * (unsigned)sh.S1) % (unsigned)sh.nTB;
                w0[i] = tb * (unsigned)sh.segT + ai;
                w1[i] = (t << MFK_W1_BBITS) | bl;
            } else {
                w0[i] = (t << MFK_W0_ABITS) | ai;
                w1[i] = ((ticket ? ticket[i] : 0u) << MFK_W1_BBITS) | bl;  // no tickets: rows are handed out by locks
            }
            slot = (long long)(hi >> sh.bitsG) * sh.nG + (long long)(hi & ((1ull << sh.bitsG) - 1ull));
            if (i > 0) {
                const unsigned long long ph = k1[i - 1] >> lowbits;
                prev = (long long)(ph >> sh.bitsG) * sh.nG + (long long)(ph & ((1ull << sh.bitsG) - 1ull));
            }
        }
        rr[i] = __uint_as_float((unsigned)val);
        for (long long s = prev + 1; s <= slot; s++) goff[s] = (unsigned)i;
        if (i == cnt - 1)
            for (long long s = slot + 1; s <= nOff; s++) goff[s] = (unsigned)cnt;
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
            int rows, int k, int k_al, float s, int zero_unseen) {
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
                val = zero_unseen ? 0.0f : __uint_as_float(0x7fc00000u);  // quiet NaN, mf/mf.cpp:996-999 (BPR: stays 0, 997)
            }
        }
        out[d] = val;
    }
    reinterpret_cast<float4 *>(dst)[0] = make_float4(out[0], out[1], out[2], out[3]);
    reinterpret_cast<float4 *>(dst)[1] = make_float4(out[4], out[5], out[6], out[7]);
}

// ------------------------------------------------------------------------------------------------
// The throughput kernel (second generation).
//
// Replaces the per-rating loop SolverBase::run + L2_MFR::prepare_for_sg_update + MFSolver::sg_update
// (mf/mf.cpp:1220-1235, 1720-1728, 1462-1548) and the block scheduler (mf/mf.cpp:113-150,193-220).
//
// Data placement: CTA c keeps S band (pass*nC + c) -- rows, AdaGrad accumulators and one ticket counter
// per row -- in shared memory for a whole pass; only the T rows and the 12-byte rating stream travel
// through L2/HBM.  A group of L lanes processes one rating (L = 8: four ratings per warp instruction).
//
// Schedule (race-free by construction, like the reference's scheduler that never co-schedules two
// blocks sharing a row band or a column band, mf/mf.cpp:130-142):
//   * T bands rotate ring-wise over the CTAs: at step t CTA c works on T band (c*S1 + t) mod nTB, and
//     group ga of the CTA owns T sub-band ga of it.  The previous owner of that sub-band is group ga of
//     CTA c+1 at step t-S1, so a group only waits for ONE counter of its neighbour CTA
//     (release/acquire in global memory); there is no grid-wide or CTA-wide barrier inside a pass.
//   * S rows are shared by the groups of one CTA through tickets: the preprocessing numbers the
//     ratings of every S row in the order (t, d, a) and a group may update the row only when the row's
//     counter equals the rating's ticket.  Every group walks its ratings in the same global order, so
//     every wait points to a strictly earlier rating: no deadlock, and the order of updates of every
//     row is fixed, so a run is reproducible bit for bit.
//   * The four groups of a warp share an instruction stream but not a queue position: a group whose
//     ticket or flag is not ready simply sits out the iteration (predicated off), it never blocks
//     its siblings.
//
// Arithmetic per rating: z by butterfly shuffle inside the group; e = r - z;
// g_t = lambda_t p - e q, g_s = lambda_s q - e p from the OLD p,q; p -= eta rsqrt(G_t) g_t;
// G += sum(g^2)/8 for BOTH halves (the shipped SSE path's rk); dims 0-7 and 8..k_al have
// separate accumulators; epoch 0 touches dims 0-7 only.
// ------------------------------------------------------------------------------------------------
// DYN: the S rows are handed out by locks instead of tickets (mfk_band_args.dynamic): whichever group asks
// first gets the row.  Still race-free (one group per row at a time) and every rating is applied exactly once,
// but the order of updates of a row depends on timing, so two runs differ in the last bits.
// prepare_for_sg_update of the six MFSolver losses, SSE code path (mf/mf.cpp:1719-1728 L2_MFR, 1768-1781 L1_MFR,
// 1829-1840 KL_MFR, 1884-1903 LR_MFC, 1965-1988 L2_MFC, 2053-2080 L1_MFC): from z = <p,q> and the rating to the scalar
// the update multiplies the other row with.  loss_add / err_add are the float terms the reference widens to double
// (err_add only differs from the loss for the two hinge losses, where it counts correctly classified ratings).
// Every operation is rounded separately (__f*_rn).  PRECISE (exact mode): exp is glibc's expf restated bit for bit
// (expf_glibc below), log -- which only feeds the loss column of the table -- goes through double precision;
// otherwise the fp32 CUDA functions.
// glibc's expf (2.27 and later: sysdeps/ieee754/flt-32/e_expf.c with the 32-entry table of e_exp2f_data.c), restated
// from its published algorithm: exp(x) = 2^(k/32) * 2^(r/32), the second factor a cubic in double precision, one final
// rounding to float.  oracle/expf_check.c compares this restatement with the C library's expf for every float with
// |x| < 87 (2.2e9 inputs): it differs for two of them (x = 32.5646324, x = -63.0994606; one ulp), far outside the
// range of a dot product of factor rows.  With it LR_MFC in exact mode is bit-exact to the reference.
__constant__ unsigned long long c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
__device__ float expf_glibc(float x) {
    if (isnan(x)) return x + x;
    if (x > 88.72283f) return __int_as_float(0x7f800000);  // overflow
    if (x < -103.972076f) return 0.0f;                     // underflow
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
                 C2 = 0x1.62e42ff0c52d6p-1 / 32, InvLn2N = 0x1.71547652b82fep+0 * 32, Shift = 0x1.8p+52;
    double z = __dmul_rn(InvLn2N, (double)x);
    double kd = __dadd_rn(z, Shift);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, Shift);
    const double r = __dsub_rn(z, kd);
    const double s = __longlong_as_double((long long)(c_exp2f_tab[ki & 31ull] + (ki << 47)));
    z = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(z, r2, y);
    return (float)__dmul_rn(y, s);
}
template <bool PRECISE>
__device__ __forceinline__ float mf_expf(float x) {
    return PRECISE ? expf_glibc(x) : expf(x);
}
template <bool PRECISE>
__device__ __forceinline__ float mf_logf(float x) {
    return PRECISE ? (float)log((double)x) : logf(x);
}
template <bool PRECISE>
__device__ __forceinline__ float loss_scalar(int fun, float z, float r, float &loss_add, float &err_add) {
    err_add = 0.f;
    switch (fun) {
        case MFK_FUN_L1_MFR: {
            z = __fsub_rn(r, z);
            loss_add = fabsf(z);
            return __fadd_rn(z > 0.f ? 1.f : 0.f, z < 0.f ? -1.f : 0.f);
        }
        case MFK_FUN_KL_MFR: {
            z = __fdiv_rn(r, z);
            loss_add = __fmul_rn(r, __fadd_rn(__fsub_rn(mf_logf<PRECISE>(z), 1.f), __fdiv_rn(1.f, z)));
            return __fsub_rn(z, 1.f);
        }
        case MFK_FUN_LR_MFC: {
            if (r > 0.f) {
                z = mf_expf<PRECISE>(-z);
                loss_add = mf_logf<PRECISE>(__fadd_rn(1.f, z));
                return __fdiv_rn(z, __fadd_rn(1.f, z));
            }
            z = mf_expf<PRECISE>(z);
            loss_add = mf_logf<PRECISE>(__fadd_rn(1.f, z));
            return __fdiv_rn(-z, __fadd_rn(1.f, z));
        }
        case MFK_FUN_L2_MFC: {
            if (r > 0.f) {
                err_add = z > 0.f ? 1.f : 0.f;
                const float t = __fsub_rn(1.f, z);
                z = 0.f > t ? 0.f : t;  // _mm_max_ps(0, t)
            } else {
                err_add = z < 0.f ? 1.f : 0.f;
                const float t = __fsub_rn(-1.f, z);
                z = 0.f < t ? 0.f : t;  // _mm_min_ps(0, t)
            }
            loss_add = __fmul_rn(z, z);
            return z;
        }
        case MFK_FUN_L1_MFC: {
            if (r > 0.f) {
                err_add = z >= 0.f ? 1.f : 0.f;
                z = __fsub_rn(1.f, z);
                loss_add = 0.f > z ? 0.f : z;
                return z >= 0.f ? 1.f : 0.f;
            }
            err_add = z < 0.f ? 1.f : 0.f;
            z = __fadd_rn(1.f, z);
            loss_add = 0.f > z ? 0.f : z;
            return z >= 0.f ? -1.f : 0.f;
        }
        default: {  // MFK_FUN_L2_MFR
            z = __fsub_rn(r, z);
            loss_add = __fmul_rn(z, z);
            return z;
        }
    }
}

// the L1 soft threshold of sg_update (mf/mf.cpp:1499-1527): sign(x) * max(|x| - step, 0), the sign taken as "x <= 0"
__device__ __forceinline__ float soft_threshold(float x, float step) {
    const unsigned flip = (x <= 0.f) ? 0x80000000u : 0u;
    float a = __fsub_rn(__uint_as_float(__float_as_uint(x) ^ flip), step);
    a = a > 0.f ? a : 0.f;
    return __uint_as_float(__float_as_uint(a) ^ flip);
}

// LATE (with DYN): the lock of an S row is taken only after the T row has arrived and is given back as soon as the new
// S row is stored -- a shorter hold time per row for a longer dependency chain per update.  Measured on B200 (one
// launch): 15 to 60 S rows per CTA (the shares of C3 when the item stripes rotate over 8 / 4 / 2 GPUs) 1.88 -> 1.55 ms,
// 3.75 -> 3.12 ms, 8.59 -> 8.12 ms; 120 rows per CTA (C3 on one GPU) 23.7 -> 24.5 ms.  The engine picks it by rows per CTA.
// GEN: any MFSolver loss (args.fun), L1 regularisation and NMF; without it the kernel is the L2_MFR fast path.
template <int L, int V, bool STATS, bool DYN, bool LATE, bool GEN>
#ifndef MFB_BAND_THREADS
#define MFB_BAND_THREADS 512
#endif
__global__ void __launch_bounds__(MFB_BAND_THREADS, 1) k_sgd_band_epoch(const __grid_constant__ mfk_band_args g) {
    // STATS: scheduling counters for tuning (MFB200_STATS=1): [0] warp iterations, [1] of them with an update,
    // [2] group updates, group-iterations without one because [3] the stream is finished, [4] the T sub-band is
    // not released yet, [5] no ticket of the window is up; [6] failed flag polls.
    unsigned long long st_[7] = {0, 0, 0, 0, 0, 0, 0};
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const mfk_band_shape &sh = g.shape;
    const int k_al = g.k_al, nvec = k_al >> 2;
    float4 *s_rows = reinterpret_cast<float4 *>(smem_raw);                          // [rows_cap][nvec]
    float2 *s_g = reinterpret_cast<float2 *>(s_rows + (size_t)sh.rows_cap * nvec);  // [rows_cap]
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_g + sh.rows_cap);              // [rows_cap]

    constexpr int GPW = 32 / L;  // groups per warp
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & (L - 1), gi = lane / L;
    const int gamma = warp * GPW + gi;
    const unsigned gmask = L == 32 ? kFull : (((1u << L) - 1u) << (gi * L));
    const bool leader = l == 0;
    const bool full = g.full != 0;  // slow_only == false
    const int nG = sh.nG;

    bool act[V], h0[V];
#pragma unroll
    for (int j = 0; j < V; j++) {
        act[j] = l + L * j < nvec;
        h0[j] = l + L * j < 2;  // dims 0-7: the first AdaGrad half
    }

    const unsigned cS1 = ((unsigned)c * (unsigned)sh.S1) % (unsigned)sh.nTB;
    unsigned *my_flag = g.flags + (size_t)c * nG + gamma;
    const unsigned *nb_flag = g.flags + (size_t)((c + 1) % sh.nC) * nG + gamma;
    double loss = 0.0, err = 0.0;
    bool dead = false;
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

        const unsigned base = g.base + (unsigned)pass * (unsigned)sh.nTB;
        const unsigned done_mark = base + (unsigned)sh.nTB;
        const unsigned pos = g.goff[(size_t)sb * nG + gamma];
        const unsigned end = g.goff[(size_t)sb * nG + gamma + 1];
        unsigned pub = base;  // value of my_flag (all earlier passes / launches are complete)

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
            if (tb >= (unsigned)sh.nTB) tb -= (unsigned)sh.nTB;
            return tb * (unsigned)sh.segT + ai;
        };
        auto pf_rows = [&](unsigned bbase, unsigned x0) {  // pull the T rows of a batch into L2
            if (bbase + (unsigned)l < end) {
                const unsigned a = t_row(x0);
                prefetch_l2_bulk(g.T + (size_t)a * k_al, (unsigned)k_al * 4u);
                prefetch_l2(g.TG + 2 * (size_t)a);
            }
        };
        constexpr unsigned kGroupBits = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
        auto low_bits = [](int nbits) -> unsigned { return nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u); };
        unsigned cbase = pos;
        ld_batch(cbase, c0, c1, cr);
        ld_batch(cbase + L, n0, n1, nr);
        ld_batch(cbase + 2 * L, m0, m1, mr);
        pf_rows(cbase, c0);
        pf_rows(cbase + L, n0);
        // The current batch is an out-of-order window: a group may take ANY pending entry of it whose ticket
        // is up and whose T row is not also the T row of an older pending entry.  The order of updates of
        // every S row (tickets) and of every T row (older-first) stays what the stream prescribes, so the
        // result does not depend on which entry is taken first.
        int nb = cbase < end ? (int)min((unsigned)L, end - cbase) : 0;  // entries in the window
        unsigned done = 0u;                                             // bit i: entry i is processed
        unsigned same;  // bit j (j < l): entry j of the window has the same T row as this lane's entry
        {
            const unsigned mt = __match_any_sync(kFull, c0);
            same = ((mt & gmask) >> (gi * L)) & low_bits(l);
        }

        int t_cur = -1;
        bool have = false;
        float4 p[V];
        float2 tg = make_float2(1.f, 1.f);
        unsigned a_row = 0, bl = 0, ticket = 0;
        int cur_idx = 0;
        float r = 0.f;
        unsigned idle = 0;
        unsigned long long idle_since = 0;

        for (;;) {
            // (1) window used up: promote the next batch (group-uniform), refresh the T-row conflict masks
            bool promoted = false;
            if (nb > 0 && done == low_bits(nb)) {
                c0 = n0; c1 = n1; cr = nr;
                n0 = m0; n1 = m1; nr = mr;
                cbase += L;
                ld_batch(cbase + 2 * L, m0, m1, mr);
                pf_rows(cbase + L, n0);
                nb = cbase < end ? (int)min((unsigned)L, end - cbase) : 0;
                done = 0u;
                promoted = true;
            }
            if (__any_sync(kFull, promoted)) {
                const unsigned mt = __match_any_sync(kFull, c0);
                same = ((mt & gmask) >> (gi * L)) & low_bits(l);
            }

            // (2) every lane judges its own window entry; the group takes the oldest eligible one
            const int myt = (int)(c0 >> MFK_W0_ABITS);
            const bool valid = l < nb && !((done >> l) & 1u);
            bool elig = false;
            if (!have && valid && myt == t_cur && !(same & ~done)) {
                const unsigned cnt = ld_acquire_cta_smem(&s_cnt[c1 & ((1u << MFK_W1_BBITS) - 1u)]);
                elig = DYN ? cnt == 0u : (cnt & MFK_TICKET_MASK) == (c1 >> MFK_W1_BBITS);
            }
            unsigned eb = (__ballot_sync(kFull, elig) >> (gi * L)) & kGroupBits;
            const unsigned vb = (__ballot_sync(kFull, valid) >> (gi * L)) & kGroupBits;
            const int sel = eb ? __ffs(eb) - 1 : (vb ? __ffs(vb) - 1 : 0);
            const unsigned x0 = __shfl_sync(kFull, c0, sel, L);
            const unsigned x1 = __shfl_sync(kFull, c1, sel, L);
            const float xr = __shfl_sync(kFull, cr, sel, L);
            if (DYN && !LATE) {  // the row looked free: try to take its lock (another group may have been faster)
                unsigned got = 0u;
                if (eb && leader) got = cas_acquire_cta_smem(&s_cnt[x1 & ((1u << MFK_W1_BBITS) - 1u)], 0u, 1u) == 0u;
                got = __shfl_sync(kFull, got, 0, L);
                __syncwarp();
                if (!got) eb = 0u;
            }

            if (eb) {  // fetch the entry: T row (from L2, prefetched a batch ago) and its accumulators
                a_row = t_row(x0);
                const float4 *trow = reinterpret_cast<const float4 *>(g.T + (size_t)a_row * k_al);
#pragma unroll
                for (int j = 0; j < V; j++)
                    p[j] = act[j] ? __ldcg(trow + l + L * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                tg = __ldcg(reinterpret_cast<const float2 *>(g.TG) + a_row);
                bl = x1 & ((1u << MFK_W1_BBITS) - 1u);
                ticket = x1 >> MFK_W1_BBITS;
                r = xr;
                cur_idx = sel;
                have = true;
            } else if (vb || (nb == 0 && pub != done_mark)) {
                // Nothing eligible.  If the oldest pending entry belongs to a later step (or the stream is finished),
                // everything of the steps before is stored.  A step may be declared complete only after ITS OWN
                // dependency has been verified -- also when the group has no rating in it -- otherwise the chain
                // "CTA c+1 released the sub-band, so CTA c+2 must have released it before" breaks at empty steps.
                const int t_new = vb ? (int)(x0 >> MFK_W0_ABITS) : sh.nTB;
                if (t_new != t_cur) {
                    // s_rel: the last step whose T sub-band the neighbour CTA has released to this group
                    // (flag >= base + s - S1 + 1); steps that precede the launch by less than S1 need nothing.
                    long long s_rel = sh.nTB;
                    if (sh.nC > 1) {
                        // acquire load (LDG.STRONG.GPU + L1 invalidate, no MEMBAR): pairs with the neighbour's
                        // fence + flag store, so this lane's T-row loads of the new step are ordered after it.
                        // Measured against relaxed polls + fence.acq_rel on success: C3 23.6 -> 22.8 ms, C2 5.61 -> 5.29
                        s_rel = (long long)(int)(ld_acquire_gpu(nb_flag) - base) + sh.S1 - 1;
                        const long long s_free = (long long)sh.S1 - 1 - (long long)pass * sh.nTB;
                        if (s_free > s_rel) s_rel = s_free;
                    }
                    // steps <= s_rel are verified, so "all steps < v complete" may be published for v <= s_rel + 1:
                    // a group with nothing to do still advances step by step, one ahead of its neighbour
                    const int v = (int)min((long long)t_new, s_rel + 1);
                    const unsigned want = base + (unsigned)v;
                    if ((int)(want - pub) > 0) {
                        fence_acq_rel_gpu();  // every lane orders its own T-row stores before the flag
                        __syncwarp(gmask);
                        if (leader) st_relaxed_gpu(my_flag, want);
                        pub = want;
                    }
                    if (vb && v == t_new && (long long)t_new <= s_rel) {
                        t_cur = t_new;
                    } else if (STATS && leader) {
                        st_[6]++;
                    }
                }
                if (STATS && leader) st_[vb ? (t_new != t_cur ? 4 : 5) : 3]++;
            } else if (STATS && leader && !have) {
                st_[3]++;
            }

            // Locks are taken late and given back early: the T row is fetched first (it belongs to this group for the
            // whole step, no lock needed), and the compare-and-swap is made to depend on the loaded registers, so a
            // group holds an S row for the arithmetic on that row alone -- not for the L2/HBM latency of its T row and
            // not for the T-side half of the update.  With few S rows per CTA (item stripes rotating over several
            // GPUs) the hold time is what bounds the launch: ratings per row x hold time.
            bool ready = have;
            if (DYN && LATE) {
                unsigned got = 0u;
                if (have && leader) {
                    unsigned x = __float_as_uint(tg.x) ^ __float_as_uint(tg.y);
#pragma unroll
                    for (int j = 0; j < V; j++) x ^= __float_as_uint(p[j].x) ^ __float_as_uint(p[j].w);
                    // any non-zero value means "held"; the data-dependent second value never matters
                    got = cas_acquire_cta_smem(&s_cnt[bl], 0u, 1u + (x == 0x7fdead01u ? 1u : 0u)) == 0u;
                }
                got = __shfl_sync(kFull, got, 0, L);
                ready = got != 0u;
                if (STATS && leader && have && !ready) st_[5]++;
            }
            if (STATS) {
                if (lane == 0) st_[0]++;
                if (ready && leader) st_[2]++;
            }
            if (!__any_sync(kFull, ready)) {
                if (__all_sync(kFull, nb == 0 && pub == done_mark)) break;
                // A wait that never ends: give up after a wall-clock limit so that the kernel terminates (and when
                // another warp or CTA has given up)
                if (++idle >= 4096u) {
                    idle = 0;
                    const unsigned long long now = global_timer_ns();
                    if (idle_since == 0) idle_since = now;
                    if (now - idle_since > g.wait_limit_ns || *reinterpret_cast<volatile int *>(g.error_flag) != 0) {
                        if (lane == 0) atomicCAS(g.error_flag, 0, 1);
                        dead = true;
                        break;
                    }
                }
                continue;
            }
            idle_since = 0;
            idle = 0;
            if (STATS && lane == 0) st_[1]++;

            // ---- the update, executed by all groups of the warp; only ready groups commit ----
            float4 *srow = s_rows + (size_t)bl * nvec;
            float4 q[V];
#pragma unroll
            for (int j = 0; j < V; j++)
                q[j] = (ready && act[j]) ? srow[l + L * j] : make_float4(0.f, 0.f, 0.f, 0.f);
            float2 sg = ready ? s_g[bl] : make_float2(1.f, 1.f);

            // z = <p,q>  (calc_z, mf/mf.cpp:1264-1273)
            float part = 0.f;
#ifdef MFB_F32X2
            // Packed fp32: every fma.rn.f32x2 (SASS FFMA2) covers two dimensions, which halves the issue slots of the
            // arithmetic -- the kernel is issue/latency-bound, not pipe-bound.
            f32x2 pp[V][2], qq[V][2];
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
#else
#pragma unroll
            for (int j = 0; j < V; j++)
                part += p[j].x * q[j].x + p[j].y * q[j].y + p[j].z * q[j].z + p[j].w * q[j].w;
#endif
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
            float e;
            if constexpr (GEN) {
                float loss_add, err_add;
                e = loss_scalar<false>(g.fun, part, r, loss_add, err_add);
                if (ready && leader) {
                    loss += (double)loss_add;
                    err += (double)err_add;
                }
            } else {
                e = r - part;  // mf/mf.cpp:1724
                if (ready && leader) loss += (double)(e * e);  // mf/mf.cpp:1725-1726
            }

#ifdef MFB_F32X2
            if constexpr (DYN && LATE && !GEN) {
                // sg_update (mf/mf.cpp:1462-1548, 1228-1234), S side first: the new S row goes to shared memory and the
                // lock is given back before anything that concerns the T row or the AdaGrad sums is computed
                const f32x2 ne2 = pack2(-e, -e);
                float ss0 = 0.f, ss1 = 0.f;
                {
                    const float eta_s0 = g.eta * rsqrtf(sg.x), eta_s1 = g.eta * rsqrtf(sg.y);
                    const f32x2 ls2 = pack2(g.lambda_s, g.lambda_s);
                    f32x2 ss1_2 = pack2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < V; j++) {
                        const float es = (j == 0 && h0[0]) ? eta_s0 : eta_s1;
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
                        if (ready && act[j] && (full || h0[j])) {
                            float4 v;
                            unpack2(qnn[0], v.x, v.y);
                            unpack2(qnn[1], v.z, v.w);
                            srow[l + L * j] = v;
                        }
                        if (j == 0) {
                            const float ss = sum2(ssj);
                            if (h0[0])
                                ss0 = ss;
                            else
                                ss1 = ss;
                        }
                    }
                    ss1 += sum2(ss1_2);
                }
                __syncwarp();  // the group's shared-memory stores are ordered before the release of the row
#ifdef MFB_RELAXED_SMEM
                if (ready && leader) st_volatile_smem(&s_cnt[bl], 0u);
#else
                if (ready && leader) st_release_cta_smem(&s_cnt[bl], 0u);
#endif
                // T side: the row belongs to this group for the whole step
                float st0 = 0.f, st1 = 0.f;
                {
                    const float eta_t0 = g.eta * rsqrtf(tg.x), eta_t1 = g.eta * rsqrtf(tg.y);
                    const f32x2 lt2 = pack2(g.lambda_t, g.lambda_t);
                    f32x2 st1_2 = pack2(0.f, 0.f);
                    float4 *trow = reinterpret_cast<float4 *>(g.T + (size_t)a_row * k_al);
#pragma unroll
                    for (int j = 0; j < V; j++) {
                        const float et = (j == 0 && h0[0]) ? eta_t0 : eta_t1;
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
                        if (ready && act[j] && (full || h0[j])) {
                            float4 v;
                            unpack2(pnn[0], v.x, v.y);
                            unpack2(pnn[1], v.z, v.w);
                            __stcg(trow + l + L * j, v);
                        }
                        if (j == 0) {
                            const float st = sum2(stj);
                            if (h0[0])
                                st0 = st;
                            else
                                st1 = st;
                        }
                    }
                    st1 += sum2(st1_2);
                }
                // AdaGrad sums: half 0 lives in lanes 0,1 of the group (chunks 0,1), half 1 everywhere else.  The T
                // accumulators are this group's; the S accumulators are added atomically (the row is no longer held).
                st0 += __shfl_xor_sync(kFull, st0, 1);
                ss0 += __shfl_xor_sync(kFull, ss0, 1);
                float2 tgn = make_float2(tg.x + st0 * 0.125f, tg.y);
                if (full) {
#pragma unroll
                    for (int o = L / 2; o > 0; o >>= 1) {
                        st1 += __shfl_xor_sync(kFull, st1, o);
                        ss1 += __shfl_xor_sync(kFull, ss1, o);
                    }
                    tgn.y += st1 * 0.125f;  // rk_slow for both halves: SURVEY.md F2
                }
                if (ready) {
                    if (leader) {
                        __stcg(reinterpret_cast<float2 *>(g.TG) + a_row, tgn);
                        atomicAdd(&s_g[bl].x, ss0 * 0.125f);
                        if (full) atomicAdd(&s_g[bl].y, ss1 * 0.125f);
                    }
                    done |= 1u << cur_idx;
                    have = false;
                }
                continue;
            }
#endif
            // sg_update for both halves (mf/mf.cpp:1462-1548, 1228-1234)
            const float eta_t0 = g.eta * rsqrtf(tg.x), eta_s0 = g.eta * rsqrtf(sg.x);
            const float eta_t1 = g.eta * rsqrtf(tg.y), eta_s1 = g.eta * rsqrtf(sg.y);
            float st0 = 0.f, ss0 = 0.f, st1 = 0.f, ss1 = 0.f;
            float4 pn[V], qn[V];
#ifdef MFB_F32X2
            {
                const f32x2 ne2 = pack2(-e, -e), lt2 = pack2(g.lambda_t, g.lambda_t), ls2 = pack2(g.lambda_s, g.lambda_s);
                f32x2 st1_2 = pack2(0.f, 0.f), ss1_2 = pack2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < V; j++) {
                    // only chunk j == 0 can belong to the first AdaGrad half (lanes 0,1 of the group)
                    const float et = (j == 0 && h0[0]) ? eta_t0 : eta_t1, es = (j == 0 && h0[0]) ? eta_s0 : eta_s1;
                    const f32x2 net2 = pack2(-et, -et), nes2 = pack2(-es, -es);
                    f32x2 stj = pack2(0.f, 0.f), ssj = pack2(0.f, 0.f);
                    f32x2 pnn[2], qnn[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const f32x2 gt = fma2(ne2, qq[j][h], mul2(lt2, pp[j][h]));
                        const f32x2 gs = fma2(ne2, pp[j][h], mul2(ls2, qq[j][h]));
                        if (j == 0) {
                            stj = fma2(gt, gt, stj);
                            ssj = fma2(gs, gs, ssj);
                        } else {
                            st1_2 = fma2(gt, gt, st1_2);
                            ss1_2 = fma2(gs, gs, ss1_2);
                        }
                        pnn[h] = fma2(net2, gt, pp[j][h]);
                        qnn[h] = fma2(nes2, gs, qq[j][h]);
                    }
                    unpack2(pnn[0], pn[j].x, pn[j].y);
                    unpack2(pnn[1], pn[j].z, pn[j].w);
                    unpack2(qnn[0], qn[j].x, qn[j].y);
                    unpack2(qnn[1], qn[j].z, qn[j].w);
                    if (j == 0) {
                        const float st = sum2(stj), ss = sum2(ssj);
                        if (h0[0]) {
                            st0 = st;
                            ss0 = ss;
                        } else {
                            st1 = st;
                            ss1 = ss;
                        }
                    }
                }
                st1 += sum2(st1_2);
                ss1 += sum2(ss1_2);
            }
#else
#pragma unroll
            for (int j = 0; j < V; j++) {
                const float et = h0[j] ? eta_t0 : eta_t1, es = h0[j] ? eta_s0 : eta_s1;
                float st = 0.f, ss = 0.f;
#define MFB_UPD(X)                                                \
    {                                                             \
        const float gt = g.lambda_t * p[j].X - e * q[j].X;        \
        const float gs = g.lambda_s * q[j].X - e * p[j].X;        \
        st += gt * gt;                                            \
        ss += gs * gs;                                            \
        pn[j].X = p[j].X - et * gt;                               \
        qn[j].X = q[j].X - es * gs;                               \
    }
                MFB_UPD(x) MFB_UPD(y) MFB_UPD(z) MFB_UPD(w)
#undef MFB_UPD
                if (h0[j]) {
                    st0 += st;
                    ss0 += ss;
                } else {
                    st1 += st;
                    ss1 += ss;
                }
            }
#endif
            if constexpr (GEN) {
                // the L1 soft threshold (mf/mf.cpp:1499-1527) and the projection on x >= 0 (1529-1541), element-wise
                // after the L2 step; the AdaGrad sums above use the L2 gradient only, as in the reference
#pragma unroll
                for (int j = 0; j < V; j++) {
                    const float et = (j == 0 && h0[0]) ? eta_t0 : eta_t1, es = (j == 0 && h0[0]) ? eta_s0 : eta_s1;
                    if (g.lambda1_t > 0.f) {
                        const float step = et * g.lambda1_t;
                        pn[j].x = soft_threshold(pn[j].x, step);
                        pn[j].y = soft_threshold(pn[j].y, step);
                        pn[j].z = soft_threshold(pn[j].z, step);
                        pn[j].w = soft_threshold(pn[j].w, step);
                    }
                    if (g.lambda1_s > 0.f) {
                        const float step = es * g.lambda1_s;
                        qn[j].x = soft_threshold(qn[j].x, step);
                        qn[j].y = soft_threshold(qn[j].y, step);
                        qn[j].z = soft_threshold(qn[j].z, step);
                        qn[j].w = soft_threshold(qn[j].w, step);
                    }
                    if (g.do_nmf) {
                        pn[j] = make_float4(fmaxf(pn[j].x, 0.f), fmaxf(pn[j].y, 0.f), fmaxf(pn[j].z, 0.f), fmaxf(pn[j].w, 0.f));
                        qn[j] = make_float4(fmaxf(qn[j].x, 0.f), fmaxf(qn[j].y, 0.f), fmaxf(qn[j].z, 0.f), fmaxf(qn[j].w, 0.f));
                    }
                }
            }
            // half 0 lives in lanes 0,1 of the group (chunks 0,1); half 1 everywhere else
            st0 += __shfl_xor_sync(kFull, st0, 1);
            ss0 += __shfl_xor_sync(kFull, ss0, 1);
            // (a group that is not ready keeps its fetched tg untouched: it will retry next iteration)
            float2 tgn = make_float2(tg.x + st0 * 0.125f, tg.y);
            sg.x += ss0 * 0.125f;
            if (full) {
#pragma unroll
                for (int o = L / 2; o > 0; o >>= 1) {
                    st1 += __shfl_xor_sync(kFull, st1, o);
                    ss1 += __shfl_xor_sync(kFull, ss1, o);
                }
                tgn.y += st1 * 0.125f;  // rk_slow for both halves: SURVEY.md F2
                sg.y += ss1 * 0.125f;
            }
            if (ready) {
                float4 *trow = reinterpret_cast<float4 *>(g.T + (size_t)a_row * k_al);
#pragma unroll
                for (int j = 0; j < V; j++)
                    if (act[j] && (full || h0[j])) {
                        srow[l + L * j] = qn[j];
                        __stcg(trow + l + L * j, pn[j]);
                    }
                if (leader) {
                    s_g[bl] = sg;
                    __stcg(reinterpret_cast<float2 *>(g.TG) + a_row, tgn);
                }
            }
            __syncwarp();  // the group's shared-memory stores are ordered before the ticket release
            if (ready) {
#ifdef MFB_RELAXED_SMEM
                if (leader) st_volatile_smem(&s_cnt[bl], DYN ? 0u : (ticket + 1u) & MFK_TICKET_MASK);
#else
                if (leader) st_release_cta_smem(&s_cnt[bl], DYN ? 0u : (ticket + 1u) & MFK_TICKET_MASK);
#endif
                done |= 1u << cur_idx;
                have = false;
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
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(kFull, loss, o);
    if (lane == 0 && loss != 0.0) atomicAdd(g.loss, loss);
    if constexpr (GEN) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(kFull, err, o);
        if (lane == 0 && err != 0.0 && g.err) atomicAdd(g.err, err);
    }
    if (STATS && g.stats) {
#pragma unroll
        for (int i = 0; i < 7; i++) {
            unsigned long long v = st_[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (lane == 0 && v) atomicAdd(g.stats + i, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The exact kernel: one thread per rating of one wavefront level; arithmetic is the SSE code path's,
// operation for operation, so results equal the reference's bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void exact_half(float *p, float *q, float *pG, float *qG, float e, int d0, int d1,
                                           float lp, float lq, float eta, const unsigned *tab, float lp1 = 0.f,
                                           float lq1 = 0.f, bool nmf = false) {
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
            // the reference's separate passes over the half (L1 threshold 1499-1527, projection 1529-1541) touch one
            // element at a time, so they can follow the L2 step element by element
            if (lp1 > 0.f) pp[j] = soft_threshold(pp[j], __fmul_rn(eta_p, lp1));
            if (lq1 > 0.f) qq[j] = soft_threshold(qq[j], __fmul_rn(eta_q, lq1));
            if (nmf) {
                pp[j] = pp[j] > 0.f ? pp[j] : 0.f;  // _mm_max_ps(x, 0)
                qq[j] = qq[j] > 0.f ? qq[j] : 0.f;
            }
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
                  float *e2_out, int fun, float lp1, float lq1, int do_nmf, float *err_out) {
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
    float loss_add, err_add;
    const float e = loss_scalar<true>(fun, z, N.r, loss_add, err_add);
    e2_out[idx] = loss_add;
    if (err_out) err_out[idx] = err_add;
    exact_half(p, q, PG + 2 * (size_t)N.u, QG + 2 * (size_t)N.v, e, 0, 8, lp, lq, eta, g_rsqrt12_table, lp1, lq1,
               do_nmf != 0);
    if (!slow_only)
        exact_half(p, q, PG + 2 * (size_t)N.u + 1, QG + 2 * (size_t)N.v + 1, e, 8, k_al, lp, lq, eta,
                   g_rsqrt12_table, lp1, lq1, do_nmf != 0);
}

// ------------------------------------------------------------------------------------------------
// One-class BPR (BPRSolver, mf/mf.cpp:2131-2335; ROW_BPR_MFOC / COL_BPR_MFOC 2608-2707), exact mode.  One thread per
// rating of a wavefront level: p (the user row; the item row when column-oriented), q (the positive), w (the negative
// the reference's scheduler drew for this rating -- computed on the host, Scheduler::get_negative 249-280).
// z = <p, q - w> in the 4-lane order of 2182-2191, the scalar exp(-z) / (1 + exp(-z)) with the float exp (glibc's,
// restated), then per half the three-row step of 2211-2323.  When w IS q the reference's loads and stores (the three
// accumulators first; p, q, w stored in this order, four dimensions at a time) make the w results win: same order here.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bpr_half(float *p, float *q, float *w, float *pG, float *qG, float *wG, float zz, int d0,
                                         int d1, float lp, float lq, float eta, const unsigned *tab, float lp1, float lq1,
                                         bool nmf) {
    const float pG0 = *pG, qG0 = *qG, wG0 = *wG;
    const float eta_p = __fmul_rn(eta, rsqrt12(pG0, tab)), eta_q = __fmul_rn(eta, rsqrt12(qG0, tab)),
                eta_w = __fmul_rn(eta, rsqrt12(wG0, tab));
    float sp[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f}, sw[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = d0; d < d1; d += 4) {
        float4 pv = *reinterpret_cast<float4 *>(p + d), qv = *reinterpret_cast<float4 *>(q + d),
               wv = *reinterpret_cast<float4 *>(w + d);
        float *pp = reinterpret_cast<float *>(&pv), *qq = reinterpret_cast<float *>(&qv), *ww = reinterpret_cast<float *>(&wv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float pg = __fadd_rn(__fmul_rn(lp, pp[j]), __fmul_rn(zz, __fsub_rn(ww[j], qq[j])));
            const float qg = __fsub_rn(__fmul_rn(lq, qq[j]), __fmul_rn(zz, pp[j]));
            const float wg = __fadd_rn(__fmul_rn(lq, ww[j]), __fmul_rn(zz, pp[j]));
            sp[j] = __fadd_rn(sp[j], __fmul_rn(pg, pg));
            sq[j] = __fadd_rn(sq[j], __fmul_rn(qg, qg));
            sw[j] = __fadd_rn(sw[j], __fmul_rn(wg, wg));
            pp[j] = __fsub_rn(pp[j], __fmul_rn(eta_p, pg));
            qq[j] = __fsub_rn(qq[j], __fmul_rn(eta_q, qg));
            ww[j] = __fsub_rn(ww[j], __fmul_rn(eta_w, wg));
        }
        *reinterpret_cast<float4 *>(p + d) = pv;
        *reinterpret_cast<float4 *>(q + d) = qv;
        *reinterpret_cast<float4 *>(w + d) = wv;
    }
    if (lp1 > 0.f) {
        const float step = __fmul_rn(eta_p, lp1);
        for (int d = d0; d < d1; d++) p[d] = soft_threshold(p[d], step);
    }
    if (lq1 > 0.f) {  // 2291-2309: q and w of four dimensions are loaded, then q is stored, then w
        const float sq1 = __fmul_rn(eta_q, lq1), sw1 = __fmul_rn(eta_w, lq1);
        for (int d = d0; d < d1; d += 4) {
            float4 qv = *reinterpret_cast<float4 *>(q + d), wv = *reinterpret_cast<float4 *>(w + d);
            qv.x = soft_threshold(qv.x, sq1); qv.y = soft_threshold(qv.y, sq1);
            qv.z = soft_threshold(qv.z, sq1); qv.w = soft_threshold(qv.w, sq1);
            wv.x = soft_threshold(wv.x, sw1); wv.y = soft_threshold(wv.y, sw1);
            wv.z = soft_threshold(wv.z, sw1); wv.w = soft_threshold(wv.w, sw1);
            *reinterpret_cast<float4 *>(q + d) = qv;
            *reinterpret_cast<float4 *>(w + d) = wv;
        }
    }
    if (nmf)
        for (int d = d0; d < d1; d++) {
            p[d] = p[d] > 0.f ? p[d] : 0.f;
            q[d] = q[d] > 0.f ? q[d] : 0.f;
            w[d] = w[d] > 0.f ? w[d] : 0.f;
        }
    *pG = __fadd_rn(pG0, __fmul_rn(__fadd_rn(__fadd_rn(sp[0], sp[1]), __fadd_rn(sp[2], sp[3])), 0.125f));
    *qG = __fadd_rn(qG0, __fmul_rn(__fadd_rn(__fadd_rn(sq[0], sq[1]), __fadd_rn(sq[2], sq[3])), 0.125f));
    *wG = __fadd_rn(wG0, __fmul_rn(__fadd_rn(__fadd_rn(sw[0], sw[1]), __fadd_rn(sw[2], sw[3])), 0.125f));
}

// order[i]: index of the rating in R; neg[i]: the negative row of this visit (an item; a user when column-oriented);
// loss_out[rating]: the float log(1 + exp(-z)) of the rating's last visit (the scheduler's per-block table, 199-200)
__global__ void __launch_bounds__(128)
k_bpr_exact_level(const mfk_node *__restrict__ R, const unsigned *__restrict__ order, const int *__restrict__ neg,
                  int first, int count, float *P, float *Q, float *PG, float *QG, int k_al, float lp, float lq, float eta,
                  int slow_only, float *loss_out, int col, float lp1, float lq1, int do_nmf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned idx = order[first + i];
    const mfk_node N = R[idx];
    const int ng = neg[first + i];
    float *p, *q, *w, *pG, *qG, *wG;
    if (!col) {
        p = P + (size_t)N.u * k_al; q = Q + (size_t)N.v * k_al; w = Q + (size_t)ng * k_al;
        pG = PG + 2 * (size_t)N.u; qG = QG + 2 * (size_t)N.v; wG = QG + 2 * (size_t)ng;
    } else {  // COL_BPR_MFOC::prepare_negative swaps the roles of the two rows (2636-2643)
        p = Q + (size_t)N.v * k_al; q = P + (size_t)N.u * k_al; w = P + (size_t)ng * k_al;
        pG = QG + 2 * (size_t)N.v; qG = PG + 2 * (size_t)N.u; wG = PG + 2 * (size_t)ng;
    }
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int d = 0; d < k_al; d += 4) {
        const float4 a = *reinterpret_cast<const float4 *>(p + d), b = *reinterpret_cast<const float4 *>(q + d),
                     c = *reinterpret_cast<const float4 *>(w + d);
        l0 = __fadd_rn(l0, __fmul_rn(a.x, __fsub_rn(b.x, c.x)));
        l1 = __fadd_rn(l1, __fmul_rn(a.y, __fsub_rn(b.y, c.y)));
        l2 = __fadd_rn(l2, __fmul_rn(a.z, __fsub_rn(b.z, c.z)));
        l3 = __fadd_rn(l3, __fmul_rn(a.w, __fsub_rn(b.w, c.w)));
    }
    float z = __fadd_rn(__fadd_rn(l0, l1), __fadd_rn(l2, l3));
    z = expf_glibc(-z);
    loss_out[idx] = (float)log((double)__fadd_rn(1.f, z));  // the float log, to the last bit in all but rare cases
    z = __fdiv_rn(z, __fadd_rn(1.f, z));
    bpr_half(p, q, w, pG, qG, wG, z, 0, 8, lp, lq, eta, g_rsqrt12_table, lp1, lq1, do_nmf != 0);
    if (!slow_only)
        bpr_half(p, q, w, pG + 1, qG + 1, wG + 1, z, 8, k_al, lp, lq, eta, g_rsqrt12_table, lp1, lq1, do_nmf != 0);
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

// calc_reg1's inner sum (mf/mf.cpp:583-606): sum_i omega_i * (float sum over all k_al dims of |row_i|), int*float
// product in float, accumulation in double.
__global__ void __launch_bounds__(128)
k_reg1(const float *__restrict__ M, const int *__restrict__ omega, int rows, int k_al, double *out) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        const int om = omega[i];
        if (om <= 0) continue;
        const float *row = M + (size_t)i * k_al;
        float t = 0.f;
        for (int d = 0; d < k_al; d++) t = __fadd_rn(t, fabsf(row[d]));
        acc += (double)__fmul_rn((float)om, t);
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

int mfk_sgd_band_max_warps(void) { return MFB_BAND_THREADS / 32; }

int mfk_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    return n;
}

int mfk_stats(const mfk_node *R, long long nnz, double *out2, void *stream) {
    k_stats<<<grid_for(nnz, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(R, nnz, out2);
    return (int)cudaGetLastError();
}

int mfk_band_keys1(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, mfk_band_shape shape,
                   float inv_scale, int *omega_p, int *omega_q, unsigned long long *keys, unsigned long long *vals,
                   unsigned long long *kept_count, int *bad_index_flag, int m, int n, mfk_hidden hidden, void *stream) {
    if (nnz <= 0) return 0;
    k_band_keys1<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        R, nnz, p_map, q_map, shape, inv_scale, omega_p, omega_q, keys, vals, kept_count, bad_index_flag, m, n, hidden);
    return (int)cudaGetLastError();
}

int mfk_owner_of(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, int swap_sides, int t_seg,
                 int world, int *omega_p, int *omega_q, unsigned char *owner, unsigned long long *counts,
                 int *bad_index_flag, int m, int n, void *stream) {
    if (nnz <= 0) return 0;
    k_owner_of<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(R, nnz, p_map, q_map, swap_sides, t_seg,
                                                                              world, omega_p, omega_q, owner, counts,
                                                                              bad_index_flag, m, n);
    return (int)cudaGetLastError();
}

// groups the ratings of a slice by destination rank: stable radix sort on the 8-bit owner with the 12-byte node as value
size_t mfk_group_tmp_bytes(long long n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned char *)nullptr, (unsigned char *)nullptr,
                                    (const mfk_node *)nullptr, (mfk_node *)nullptr, n, 0, 8);
    return b;
}
int mfk_group_by_owner(const unsigned char *owner_in, unsigned char *owner_out, const mfk_node *nodes_in,
                       mfk_node *nodes_out, long long n, int owner_bits, void *tmp, size_t tmp_bytes, void *stream) {
    return (int)cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, owner_in, owner_out, nodes_in, nodes_out, n, 0, owner_bits,
                                                (cudaStream_t)stream);
}

size_t mfk_sort_tmp_bytes(long long n) {
    size_t b32 = 0, b64 = 0, bsc = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b32, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const unsigned *)nullptr, (unsigned *)nullptr, n, 0, 64);
    cub::DeviceRadixSort::SortPairs(nullptr, b64, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const unsigned long long *)nullptr, (unsigned long long *)nullptr, n, 0, 64);
    cub::DeviceScan::InclusiveScan(nullptr, bsc, (const unsigned *)nullptr, (unsigned *)nullptr, MaxU32(), n);
    return std::max(bsc, std::max(b32, b64));
}

int mfk_sort_pairs32(unsigned long long *keys_in, unsigned long long *keys_out, unsigned *vals_in,
                     unsigned *vals_out, long long n, int end_bit, void *tmp, size_t tmp_bytes, void *stream) {
    return (int)cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned long long *)keys_in, keys_out,
                                                (const unsigned *)vals_in, vals_out, n, 0, end_bit,
                                                (cudaStream_t)stream);
}

int mfk_sort_pairs64(unsigned long long *keys_in, unsigned long long *keys_out, unsigned long long *vals_in,
                     unsigned long long *vals_out, long long n, int end_bit, void *tmp, size_t tmp_bytes,
                     void *stream) {
    return (int)cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned long long *)keys_in, keys_out,
                                                (const unsigned long long *)vals_in, vals_out, n, 0, end_bit,
                                                (cudaStream_t)stream);
}

int mfk_band_segstart(const unsigned long long *keys1_sorted, long long nnz, mfk_band_shape shape, unsigned *head,
                      unsigned *segstart, void *tmp, size_t tmp_bytes, void *stream) {
    if (nnz <= 0) return 0;
    k_band_seghead<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(keys1_sorted, nnz, shape, head);
    return (int)cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, (const unsigned *)head, segstart, MaxU32(), nnz,
                                               (cudaStream_t)stream);
}

int mfk_band_rank_bits(mfk_band_shape shape) {
    const int r = 64 - shape.bitsB - shape.bitsT - shape.bitsG;
    return r > 24 ? 24 : r;
}

int mfk_band_keys2(const unsigned long long *keys1_sorted, const unsigned long long *vals1_sorted,
                   const unsigned *segstart, long long nnz, mfk_band_shape shape, unsigned long long *keys2,
                   unsigned *idx, void *stream) {
    if (nnz <= 0) return 0;
    k_band_keys2<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        keys1_sorted, vals1_sorted, segstart, nnz, shape, mfk_band_rank_bits(shape), keys2, idx);
    return (int)cudaGetLastError();
}

int mfk_band_tickets(const unsigned long long *keys2_sorted, const unsigned *idx_sorted, long long nnz,
                     mfk_band_shape shape, unsigned *first, unsigned *ticket, void *stream) {
    if (nnz <= 0) return 0;
    const int shift = shape.bitsT + mfk_band_rank_bits(shape) + shape.bitsG;
    k_band_heads<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(keys2_sorted, nnz, shift, first);
    k_band_tickets<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(keys2_sorted, idx_sorted, nnz, shift,
                                                                                  first, ticket);
    return (int)cudaGetLastError();
}

int mfk_band_stream(const unsigned long long *keys1_sorted, const unsigned long long *vals1_sorted,
                    const unsigned *ticket, long long nnz, mfk_band_shape shape, unsigned *w0, unsigned *w1, float *rr,
                    unsigned *goff, void *stream) {
    if (nnz <= 0) return 0;
    k_band_stream<<<grid_for(nnz, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(keys1_sorted, vals1_sorted, ticket,
                                                                                 nnz, shape, w0, w1, rr, goff);
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
                  int zero_unseen, void *stream) {
    const long long threads = (long long)rows * (k_al / 8);
    if (threads == 0) return 0;
    const float s = (float)sqrt(1.0 / (double)k);  // mf/mf.cpp:971
    k_init_rows<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, G, omega, rank, rank_base,
                                                                                    rows, k, k_al, s, zero_unseen);
    return (int)cudaGetLastError();
}

int mfk_sgd_band_max_smem(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    return v;
}

int mfk_sgd_band_epoch(const mfk_band_args *args, void *stream) {
    const int nvec = args->k_al / 4;
    const int L = args->shape.L;
    const void *fn = nullptr;
    const bool st = args->stats != nullptr;
    const bool dy = args->dynamic != 0;
    const bool lt = dy && args->late_lock != 0;
    // the general update: no counters, no late lock (dispatch only; the fast path stays untouched)
    const bool gn = args->fun != MFK_FUN_L2_MFR || args->lambda1_s > 0.f || args->lambda1_t > 0.f || args->do_nmf != 0;
#define MFB_PICK(LL, VV)                                                                                            \
    (gn ? (dy ? (const void *)k_sgd_band_epoch<LL, VV, false, true, false, true>                                    \
              : (const void *)k_sgd_band_epoch<LL, VV, false, false, false, true>)                                  \
        : st ? (dy ? (lt ? (const void *)k_sgd_band_epoch<LL, VV, true, true, true, false>                          \
                         : (const void *)k_sgd_band_epoch<LL, VV, true, true, false, false>)                        \
                   : (const void *)k_sgd_band_epoch<LL, VV, true, false, false, false>)                             \
             : (dy ? (lt ? (const void *)k_sgd_band_epoch<LL, VV, false, true, true, false>                         \
                         : (const void *)k_sgd_band_epoch<LL, VV, false, true, false, false>)                       \
                   : (const void *)k_sgd_band_epoch<LL, VV, false, false, false, false>))
    if (L == 8) {
        const int v = (nvec + 7) / 8;
        if (v <= 1) fn = MFB_PICK(8, 1);
        else if (v == 2) fn = MFB_PICK(8, 2);
        else if (v == 3) fn = MFB_PICK(8, 3);
        else if (v == 4) fn = MFB_PICK(8, 4);
    } else if (L == 16) {
        const int v = (nvec + 15) / 16;
        if (v <= 1) fn = MFB_PICK(16, 1);
        else if (v == 2) fn = MFB_PICK(16, 2);
    } else if (L == 32) {
        const int v = (nvec + 31) / 32;
        if (v <= 1) fn = MFB_PICK(32, 1);
        else if (v <= 2) fn = MFB_PICK(32, 2);
        else if (v <= 4) fn = MFB_PICK(32, 4);
    }
#undef MFB_PICK
    if (!fn) return (int)cudaErrorInvalidValue;  // k > 512 is not supported by the band kernel
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)args->shape.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    void *kargs[] = {(void *)args};
    dim3 grid(args->shape.nC), block(args->shape.nWarps * 32);
    return (int)cudaLaunchCooperativeKernel(fn, grid, block, kargs, args->shape.smem_bytes, (cudaStream_t)stream);
}

int mfk_sgd_exact_level(const mfk_node *R, const unsigned *order, int count, float *P, float *Q, float *PG,
                        float *QG, int k_al, float lambda_p, float lambda_q, float eta, int slow_only,
                        float *e2_out, int fun, float lambda_p1, float lambda_q1, int do_nmf, float *err_out,
                        void *stream) {
    if (count <= 0) return 0;
    int rc = ensure_table();
    if (rc) return rc;
    k_sgd_exact_level<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        R, order, count, P, Q, PG, QG, k_al, lambda_p, lambda_q, eta, slow_only, e2_out, fun, lambda_p1, lambda_q1,
        do_nmf, err_out);
    return (int)cudaGetLastError();
}

int mfk_bpr_exact_level(const mfk_node *R, const unsigned *order, const int *neg, int first, int count, float *P, float *Q,
                        float *PG, float *QG, int k_al, float lambda_p, float lambda_q, float eta, int slow_only,
                        float *loss_out, int col_oriented, float lambda_p1, float lambda_q1, int do_nmf, void *stream) {
    if (count <= 0) return 0;
    const int trc = ensure_table();
    if (trc) return trc;
    k_bpr_exact_level<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(R, order, neg, first, count, P, Q, PG, QG, k_al,
                                                                           lambda_p, lambda_q, eta, slow_only, loss_out,
                                                                           col_oriented, lambda_p1, lambda_q1, do_nmf);
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

int mfk_reg1(const float *M, const int *omega, int rows, int k_al, double *out1, void *stream) {
    if (rows <= 0) return 0;
    k_reg1<<<grid_for(rows, 128, 148 * 8), 128, 0, (cudaStream_t)stream>>>(M, omega, rows, k_al, out1);
    return (int)cudaGetLastError();
}

int mfk_finalize_rows(const float *M, const int *map, int rows, int k, int k_al, float factor, float *out,
                      void *stream) {
    if ((long long)rows * k == 0) return 0;
    k_finalize_rows<<<grid_for((long long)rows * k, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        M, map, rows, k, k_al, factor, out);
    return (int)cudaGetLastError();
}

}  // extern "C"
