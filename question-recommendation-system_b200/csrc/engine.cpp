// csrc/engine.cpp -- host orchestration of the training path.  Replaces fpsg + fpsg_core
// (mf/mf.cpp:2945-3042, 2774-2943) of the reference; all arithmetic on ratings and factors happens
// in the kernels behind kernels.h.  There is no CPU fallback: without a CUDA device load() fails.
//
// Two modes:
//   EXACT  the reference's single-thread update order (its only reproducible mode),
//          replayed as wavefronts: ratings whose rows and columns are disjoint commute exactly, so
//          each wavefront is one parallel launch and the result equals the sequential one bit for
//          bit.  The host only orders the ratings (grid, block schedule, levels).
//   RING   the throughput schedule (see kernels.cu, k_sgd_band_epoch): the smaller factor matrix lives
//          in shared memory band by band, the other one streams ring-wise; preprocessing on the device.
#include "engine.hpp"
#include "nccl_dl.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <mutex>

namespace mfb200 {

namespace {
thread_local std::string t_error;

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int env_int(const char *name, int dflt) {
    const char *s = std::getenv(name);
    return (s && *s) ? std::atoi(s) : dflt;
}
int ceil_div(int a, int b) { return (a + b - 1) / b; }
struct Trace {  // MFB200_TRACE=1: phase timings on stderr
    bool on;
    double t;
    Trace() : on(env_int("MFB200_TRACE", 0) != 0), t(now_ms()) {}
    void mark(const char *what) {
        if (!on) return;
        cudaDeviceSynchronize();
        const double n = now_ms();
        std::fprintf(stderr, "mfb200 trace: %-28s %9.2f ms\n", what, n - t);
        t = n;
    }
};
int bits_for(long long v) {  // number of bits needed to represent values in [0, v)
    int b = 1;
    while ((1ll << b) < v) b++;
    return b;
}
}  // namespace

void set_error(const std::string &msg) {
    t_error = msg;
    std::cerr << "mfb200: " << msg << std::endl;
}
const char *last_error() { return t_error.c_str(); }

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e__ = (cudaError_t)(call);                                                      \
        if (e__ != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                      std::to_string(__LINE__) + ")");                                              \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

// Device memory comes from the device's stream-ordered pool with an unlimited release threshold: a second
// training call in the same process (the PHP worker case) gets its gigabytes back from the pool instead of
// paying cudaMalloc/cudaFree again (measured ~0.3 s per call at 100M ratings).
static thread_local cudaStream_t t_pool_stream = nullptr;
static int pool_setup(int device) {
    static bool done[64] = {false};
    if (device >= 0 && device < 64 && done[device]) return 0;
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    if (device >= 0 && device < 64) done[device] = true;
    return 0;
}
// The small pinned block a session reads its loss sums through.  cudaFreeHost synchronises the device and was measured at
// 1 ... 380 ms per call (the whole spread between two "identical" end-to-end calls of bench.py), so the blocks are kept
// for the process: a session takes one from this list and gives it back.
namespace {
struct PinnedBlock {
    void *p;
    size_t bytes;
};
std::mutex g_pinned_mu;
std::vector<PinnedBlock> g_pinned_free;   // blocks nobody uses at the moment
std::vector<PinnedBlock> g_pinned_lent;   // blocks a session holds (to know their size when they come back)
}  // namespace
// smallest free block of at least `bytes`, or a new one (sizes rounded up to a power of two: blocks get reused)
static int pinned_take(void **out, size_t bytes) {
    size_t want = 4096;
    while (want < bytes) want <<= 1;
    {
        std::lock_guard<std::mutex> lock(g_pinned_mu);
        int best = -1;
        for (int i = 0; i < (int)g_pinned_free.size(); i++)
            if (g_pinned_free[(size_t)i].bytes >= want && (best < 0 || g_pinned_free[(size_t)i].bytes < g_pinned_free[(size_t)best].bytes))
                best = i;
        if (best >= 0) {
            g_pinned_lent.push_back(g_pinned_free[(size_t)best]);
            *out = g_pinned_free[(size_t)best].p;
            g_pinned_free.erase(g_pinned_free.begin() + best);
            return 0;
        }
    }
    void *p = nullptr;
    CK(cudaMallocHost(&p, want));
    std::lock_guard<std::mutex> lock(g_pinned_mu);
    g_pinned_lent.push_back(PinnedBlock{p, want});
    *out = p;
    return 0;
}
static void pinned_give(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pinned_mu);
    for (size_t i = 0; i < g_pinned_lent.size(); i++)
        if (g_pinned_lent[i].p == p) {
            g_pinned_free.push_back(g_pinned_lent[i]);
            g_pinned_lent.erase(g_pinned_lent.begin() + (long)i);
            return;
        }
}
static int pinned_acc_take(double **p, size_t doubles) { return pinned_take((void **)p, sizeof(double) * doubles); }
static void pinned_acc_give(double *p) { pinned_give(p); }

template <typename T>
static int dev_alloc(T **p, size_t count) {
    *p = nullptr;
    CK(cudaMallocAsync((void **)p, sizeof(T) * (count ? count : 1), t_pool_stream));
    return 0;
}
template <typename T>
static void dev_free(T *&p) {
    if (p) cudaFreeAsync(p, t_pool_stream);
    p = nullptr;
}

// ------------------------------------------------------------------------------------------------
// Host <-> device copies of pageable buffers (the caller's rating array, the caller's factor arrays), pipelined
// through two pinned staging buffers: the CPU copies chunk i+1 (several threads) while the DMA moves chunk i.
// Measured at 1.2 GB: 106 ms with a plain cudaMemcpyAsync from pageable memory.
namespace {
struct Staging {
    static constexpr size_t kChunk = 32u << 20;
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool ok = false, tried = false;
    std::mutex mu;  // one copy at a time per device: the two buffers are shared by every caller of that device
    void init() {
        if (tried) return;
        tried = true;
        ok = cudaMallocHost(&buf[0], kChunk) == cudaSuccess && cudaMallocHost(&buf[1], kChunk) == cudaSuccess &&
             cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming) == cudaSuccess;
    }
};
// One pair of pinned buffers per device (process lifetime, like the CUDA context): the events belong to the device
// that was current when they were created, and several devices (MFB200_GPUS) copy at the same time.
Staging &staging() {
    static Staging s[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return s[dev >= 0 && dev < 64 ? dev : 0];
}
void par_memcpy(void *dst, const void *src, size_t bytes) {
    static const int nt = std::max(1, std::min(env_int("MFB200_COPY_THREADS", 8), 64));
    const size_t part = (bytes / nt + 63) & ~(size_t)63;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; t++) {
        const size_t lo = std::min(bytes, part * t), hi = t == nt - 1 ? bytes : std::min(bytes, part * (t + 1));
        if (hi > lo) std::memcpy((char *)dst + lo, (const char *)src + lo, hi - lo);
    }
}
}  // namespace

// A staging buffer is written (by the CPU here, by a D2H copy in staged_d2h) only after the DMA that last used it
// has completed -- whichever call and whichever stream issued it: done[i] is waited for before EVERY reuse
// (synchronising an event that was never recorded returns at once).  Round 1 waited only from the third chunk of a
// call on, so the first two chunks of a call could overwrite the tail of the previous call's upload in flight
// (three >= 128 MB uploads back to back, config #4: wrong factors on the device).
static int staged_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st) {
    Staging &sg = staging();
    std::lock_guard<std::mutex> lock(sg.mu);
    sg.init();
    if (!sg.ok || bytes < 4 * Staging::kChunk) {
        CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    int i = 0;
    for (size_t off = 0; off < bytes; off += Staging::kChunk, i ^= 1) {
        const size_t n = std::min(Staging::kChunk, bytes - off);
        CK(cudaEventSynchronize(sg.done[i]));  // the buffer's previous chunk (of this or of an earlier call) has left
        par_memcpy(sg.buf[i], (const char *)src_host + off, n);
        CK(cudaMemcpyAsync((char *)dst_dev + off, sg.buf[i], n, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(sg.done[i], st));
    }
    return 0;
}

// synchronous with respect to the host: dst_host is complete on return
static int staged_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st) {
    Staging &sg = staging();
    std::lock_guard<std::mutex> lock(sg.mu);
    sg.init();
    if (!sg.ok || bytes < 4 * Staging::kChunk) {
        CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    const size_t nchunks = (bytes + Staging::kChunk - 1) / Staging::kChunk;
    for (size_t c = 0; c < nchunks + 1; c++) {
        if (c < nchunks) {
            const size_t off = c * Staging::kChunk, n = std::min(Staging::kChunk, bytes - off);
            // (c >= 2: the CPU copy out of this buffer finished in iteration c-1; c < 2: an earlier call's upload
            // from this buffer, possibly on another stream, must have left)
            CK(cudaEventSynchronize(sg.done[c & 1]));
            CK(cudaMemcpyAsync(sg.buf[c & 1], (const char *)src_dev + off, n, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(sg.done[c & 1], st));
        }
        if (c > 0) {  // while chunk c is on the bus, chunk c-1 goes from the staging buffer to the caller's array
            const size_t off = (c - 1) * Staging::kChunk, n = std::min(Staging::kChunk, bytes - off);
            CK(cudaEventSynchronize(sg.done[(c - 1) & 1]));
            par_memcpy((char *)dst_host + off, sg.buf[(c - 1) & 1], n);
        }
    }
    return 0;
}

// The same two services for the one-shot calls of mf_api.cpp (predict, metrics, top-k): pooled device memory on the
// legacy stream and staged copies of the caller's pageable arrays.
int api_pool_alloc(void **p, size_t bytes) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (pool_setup(dev)) return 1;
    *p = nullptr;
    CK(cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t) nullptr));
    return 0;
}
void api_pool_free(void *p) {
    if (p) cudaFreeAsync(p, (cudaStream_t) nullptr);
}
int api_h2d(void *dst_dev, const void *src_host, size_t bytes) { return staged_h2d(dst_dev, src_host, bytes, nullptr); }
int api_d2h(void *dst_host, const void *src_dev, size_t bytes) { return staged_d2h(dst_host, src_dev, bytes, nullptr); }

// ------------------------------------------------------------------------------------------------
// Shape of the band schedule (kernels.h, mfk_band_shape).  S = the side with fewer rows: it is what a
// CTA keeps in shared memory, so its bands must fit there; T = the other side, which streams.
// kernel: 0 = band kernel (kernels.cu), 1 = run kernel (sgd_run.cu), 2 = cell kernel (sgd_cell.cu), 3 = warp kernel
// (sgd_warp.cu), 4 = run or warp, whichever suits the size of a launch (cell and warp need locks: the caller asks for
// 1 when the run must be reproducible)
bool plan_band(int m, int n, long long nnz, int k_al, int sm_count, int max_smem, int world, int rank,
               mfk_band_shape *out, int kernel) {
    mfk_band_shape s;
    std::memset(&s, 0, sizeof(s));
    s.swap_sides = n > m ? 1 : 0;
    const int nS = std::min(m, n), nT = std::max(m, n);
    world = std::max(1, world);
    // stripes per rank: 1 = a transfer sits between two launches (a few tens of microseconds); 2 = half-stripes,
    // the transfer overlaps the next launch, but every launch sees half as many S rows per CTA
    const int spr = std::max(1, std::min(env_int("MFB200_STRIPES_PER_RANK", 1), 4));
    s.nStripes = world > 1 ? spr * world : 1;
    s.stripeRows = std::max(1, ceil_div(nS, s.nStripes));
    s.tSeg = std::max(1, ceil_div(nT, world));  // T rows per rank
    s.tLo = std::min(nT, rank * s.tSeg);
    s.tRows = std::max(0, std::min(nT, s.tLo + s.tSeg) - s.tLo);
    const long long nnz_launch = std::max<long long>(1, nnz / ((long long)s.nStripes * world));

    s.L = k_al <= 128 ? 8 : 32;
    {
        const int want = env_int("MFB200_GROUP_LANES", 0);  // tuning: 8, 16 or 32 lanes per rating
        if ((want == 16 && k_al <= 128) || want == 32) s.L = want;
    }
    const bool can_row = s.L == 8 && k_al <= 128;
    if (!can_row) kernel = 0;
    // warps per CTA: what the kernel was compiled for (register budget); the run kernel fits more than the band kernel
    const bool run_kind = kernel == 1 || kernel == 4 || kernel == 5 || kernel == 6;
    const int max_warps = run_kind ? mfk_sgd_run_max_warps() : mfk_sgd_band_max_warps();
    // (run kernel: 20 warps when a launch has enough ratings per cell to keep 80 groups busy, measured break-even between
    // 4.5 and 14 ratings per (group, step) cell; never with T-row locks, which are for the small launches)
    const double cell16 = (double)nnz_launch / ((double)std::min(sm_count, std::max(1, std::min(s.stripeRows, s.tRows))) *
                                                std::min(sm_count, std::max(1, std::min(s.stripeRows, s.tRows))) * 64.0);
    const int dflt_warps = run_kind && cell16 >= (double)env_int("MFB200_W20_ABOVE", 10) ? 20 : 16;
    s.nWarps = std::max(1, std::min(env_int("MFB200_RING_WARPS", dflt_warps), max_warps));
    s.nG = s.nWarps * 32 / s.L;
    s.S1 = 1;  // decided below, once the number of CTAs is known
    const int max_ctas = std::min(sm_count, std::min(s.stripeRows, std::max(1, s.tRows)));
    // CTAs: one per SM, but never so many that a (step, group) cell holds less than ~min_cell ratings
    const int min_cell = std::max(1, env_int("MFB200_MIN_CELL", 4));
    int nC = max_ctas;
    const long long by_work = (long long)std::floor(std::sqrt((double)nnz_launch / ((double)min_cell * s.S1 * s.nG)));
    nC = (int)std::max<long long>(1, std::min<long long>(nC, by_work));
    if (kernel == 4) {
        // Few ratings per (group, step) cell -- item stripes rotating over 4 or 8 GPUs, config #1: the run kernel with
        // T-row locks on every SM instead of the ring hand-off on fewer CTAs (measured per launch: 0.95 against 1.18 ms
        // for one rank's share of config #3 on 8 GPUs, 2.30 against 2.46 on 4, 0.83 against 0.91 for config #1; slower
        // from ~15 ratings per cell on: profiles/experiments/r2_tlock_kernel.txt).  The warp kernel is never picked by
        // itself (slower than the run kernel at every size; MFB200_KERNEL=warp selects it).
        const double cell = (double)nnz_launch / ((double)max_ctas * max_ctas * s.nG);
        kernel = cell < (double)env_int("MFB200_WARP_BELOW", 0) ? 3 : cell < (double)env_int("MFB200_TLOCK_BELOW", 6) ? 5 : 1;
        // The item kernel where it fits exactly: about one S row per group of a CTA (64 groups), so that every group owns
        // one row for the launch and nothing is ever handed over -- C3's share on 2 GPUs, 60 item rows per CTA: 5.93 against
        // 6.29 ms.  With two rows per group (C3 on one GPU) or a fraction (4, 8 GPUs) it loses (profiles/experiments).
        const double rows_per_cta = (double)s.stripeRows / (double)max_ctas;
        if (kernel == 1 && env_int("MFB200_ITEM_AUTO", 1) && rows_per_cta >= 48.0 && rows_per_cta <= 66.0 && s.stripeRows >= 64 * 16)
            kernel = 6;
    }
    if (kernel == 5) {  // the run kernel with T-row locks: no step hand-off, so no lower bound on the ratings of a cell
        s.tlock = 1;
        kernel = 1;
        nC = max_ctas;
        s.nWarps = std::min(s.nWarps, 16);  // (+ the releaser warp: the 17-warp build has no spills, a 21-warp one would)
        s.nG = s.nWarps * 32 / s.L;
    }
    if (kernel == 6) {  // the item kernel: every S row belongs to one group for the launch, T rows by locks, one pass
        s.tlock = 1;
        nC = max_ctas;
        s.L = env_int("MFB200_ITEM_LANES", 8) == 32 ? 32 : 8;
        s.nWarps = std::max(1, std::min(env_int("MFB200_RING_WARPS", mfk_sgd_item_max_warps(s.L)), mfk_sgd_item_max_warps(s.L)));
        s.nG = s.nWarps * 32 / s.L;
    }
    if (kernel == 3) {  // the warps own the T sub-bands: 4x fewer, 4x larger cells
        s.nG = s.nWarps;
        const long long bw = (long long)std::floor(std::sqrt((double)nnz_launch / ((double)min_cell * s.S1 * s.nG)));
        nC = (int)std::max<long long>(1, std::min<long long>(max_ctas, bw));
    }
    if (kernel == 2) {
        // A T band visits every CTA in turn, and every visit ends with a hand-off (fence, flag through L2, poll, fence,
        // first T-row load: a few microseconds), so a launch lasts at least nC * S1/(S1-1) hand-offs; the work per CTA
        // falls with nC.  The two meet at nC ~ sqrt(coef * ratings per launch).
        const double coef = (double)env_int("MFB200_CELL_COEF_E4", 68) * 1e-4;
        nC = (int)std::max(16.0, std::ceil(std::sqrt(coef * (double)nnz_launch)));
        nC = std::min(nC, max_ctas);
    }
    nC = std::max(1, std::min(env_int("MFB200_RING_CTAS", nC), sm_count));
    nC = std::min(nC, std::min(s.stripeRows, std::max(1, s.tRows)));
    s.nC = nC;
    if (kernel == 2) {
        // fine steps cost nothing here (no group walks them): eight T bands per CTA hide the hand-off behind seven
        // steps of work; a T band keeps at least a few rows
        s.S1 = std::max(1, std::min(env_int("MFB200_CELL_S1", 8), std::max(1, s.tRows / (4 * nC))));
        s.S1 = std::min(s.S1, 64);
    }
    const int row_bytes = k_al * 4 + 12;  // row + two accumulators + ticket counter
    // the run kernel (sgd_run.cu) also keeps one prefetch slot per group in shared memory and wants the ratings of a T
    // row adjacent in the stream; the cell kernel adds one counter per step
    s.by_row = kernel == 6 ? 4 : kernel;
    if (kernel == 6) {
        s.nC = nC;
        s.S1 = 1;
        s.nPass = 1;
        s.segS = std::max(1, ceil_div(s.stripeRows, s.nC));
        s.nTB = std::max(1, ceil_div(s.segS, s.nG));  // the step field of the key orders a group's S rows
        s.rows_cap = s.segS;
        s.smem_bytes = mfk_sgd_item_smem_bytes(k_al, s.nG);
        s.segT = std::max(1, s.tRows);
        s.segT2 = s.segT;
        s.bitsA = bits_for(s.segT);
        s.bitsT = bits_for(s.nTB);
        s.bitsD = bits_for(s.nG);
        s.bitsG = s.bitsD;
        s.bitsSB = bits_for((long long)s.nStripes * s.nC * s.nPass);
        s.bitsB = bits_for(nS);
        if (s.segS >= (1 << MFK_W1_BBITS) || s.bitsA > (int)MFK_W0_ABITS || s.bitsSB + s.bitsG + s.bitsT + s.bitsD + s.bitsA + 1 > 64) {
            set_error("problem shape does not fit the item kernel's key encoding");
            return false;
        }
        *out = s;
        return true;
    }
    const int slot_bytes = kernel == 2 ? (int)mfk_sgd_cell_extra_bytes(k_al, s.nG, nC * s.S1)
                           : kernel == 1 ? (int)mfk_sgd_run_slot_bytes(k_al, s.nG)
                           : kernel == 3 ? (int)mfk_sgd_run_slot_bytes(k_al, 4 * s.nWarps) : 0;
    const int cap = std::min((1 << MFK_W1_BBITS) - 1, (max_smem - 1024 - slot_bytes) / row_bytes);
    if (cap < 1) {
        set_error("a factor row does not fit in shared memory");
        return false;
    }
    const int per_cta = ceil_div(s.stripeRows, s.nC);
    s.nPass = std::max(1, ceil_div(per_cta, cap));
    s.segS = std::max(1, ceil_div(s.stripeRows, s.nC * s.nPass));
    // Slack: with two T bands per CTA a group waits for its neighbour's step t-2 instead of t-1, which removes most
    // hand-off waits, but halves the ratings per (group, step) cell.  Measured: pays off at ~70 ratings per cell
    // (Netflix shape, -6 %), costs at ~14 (MovieLens shape, +9 %).
    if (kernel != 2) {
        const double cell = (double)nnz_launch / ((double)nC * nC * s.nG * s.nPass);
        s.S1 = cell >= 48.0 ? 2 : 1;
        // (the run kernel at 20 warps in one pass: two T bands per CTA already pay at ~11 ratings per cell -- C2 4.49 -> 4.38 ms,
        // C3's share on 2 GPUs 6.39 -> 6.29; with several passes, config #4, they do not)
        if (kernel == 1 && s.nWarps == 20 && s.nPass == 1 && cell >= 10.0) s.S1 = 2;
        s.S1 = std::max(1, std::min(env_int("MFB200_RING_S1", s.S1), 16));
    }
    s.nTB = s.nC * s.S1;
    if (kernel == 2) {
        // entries a group claims at a time: about half of a group's fair share of a cell, at most one per lane
        const double cell = (double)nnz_launch / ((double)nC * s.nTB * s.nPass);
        int ch = 1;
        while (ch < 8 && (double)(2 * ch) * 2.0 * s.nG <= cell) ch *= 2;
        s.chunk = std::max(1, std::min(env_int("MFB200_CELL_CHUNK", ch), 8));
    }

    s.rows_cap = s.segS;
    s.smem_bytes = (unsigned)s.segS * (unsigned)row_bytes + (unsigned)slot_bytes;
    s.segT = std::max(1, ceil_div(std::max(1, s.tRows), s.nTB));
    s.segT2 = std::max(1, ceil_div(s.segT, s.nG));
    s.bitsA = bits_for(s.segT);
    s.bitsT = bits_for(s.nTB);
    s.bitsD = kernel == 2 ? 0 : bits_for(s.nG);
    s.bitsG = s.bitsD;
    s.bitsSB = bits_for((long long)s.nStripes * s.nC * s.nPass);
    s.bitsB = bits_for(nS);
    if (s.bitsA > (int)MFK_W0_ABITS || s.bitsT > 32 - (int)MFK_W0_ABITS ||
        s.bitsB + s.bitsT + s.bitsG + 8 > 64 || s.bitsSB + s.bitsG + s.bitsT + s.bitsD + s.bitsA + 1 > 64 ||
        (kernel == 2 && (long long)s.tRows > (long long)MFK_CELL_ROW_MASK)) {
        set_error("problem shape does not fit the band schedule's key encoding");
        return false;
    }
    *out = s;
    return true;
}

// The all-to-all of the sharded load: counts[q * world + d] = ratings rank q holds for destination d (every rank has
// the whole matrix after an all-gather).  Rank `me` sends its block for d from offset send_off[d] of its grouped slice
// and receives rank q's block at recv_off[q]: blocks land in rank order.  Returns the number of ratings received.
long long exchange_plan(int world, int me, const unsigned long long *counts, long long *send_off, long long *recv_off) {
    long long so = 0, ro = 0;
    for (int q = 0; q < world; q++) {
        send_off[q] = so;
        recv_off[q] = ro;
        so += (long long)counts[(size_t)me * world + q];
        ro += (long long)counts[(size_t)q * world + me];
    }
    return ro;
}
// owner of a T row in the sharded load (k_owner_of uses the same expression on the device)
int owner_of_row(int t_row, int t_seg, int world) { return std::min(t_row / std::max(1, t_seg), world - 1); }

RotationStep rotation_step(int world, int rank, long long substep, int spr) {
    RotationStep r;
    const int ns = spr * world, sigma = (int)(substep % ns);
    r.compute = (spr * rank + sigma) % ns;
    r.send_stripe = r.compute;
    r.send_to = (rank + world - 1) % world;
    r.recv_from = (rank + 1) % world;
    r.recv_stripe = (spr * r.recv_from + sigma) % ns;
    return r;
}

// ------------------------------------------------------------------------------------------------
Session::Session(int m, int n, const mfb200_param &prm, int rank, int world, const void *nccl_id)
    : m_(m), n_(n), prm_(prm) {
    k_ = prm.k;
    k_al_ = ((k_ + 7) / 8) * 8;  // mf/mf.cpp:959
    rank_ = rank;
    world_ = std::max(1, world);
    std::memset(nccl_id_, 0, sizeof(nccl_id_));
    if (nccl_id) std::memcpy(nccl_id_, nccl_id, sizeof(nccl_id_));
}

#define NCK(call)                                                                                       \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            set_error(std::string(#call) + ": " + nccl_api()->GetErrorString(r__) + " (" + __FILE__ + ":" + \
                      std::to_string(__LINE__) + ")");                                                  \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)

Session::~Session() { free_all(); }

void Session::free_all() {
    if (!device_ready_ && !stream_) return;  // nothing was ever created, or release() has already run
    if (device_ready_) cudaSetDevice(device_);
    t_pool_stream = (cudaStream_t)stream_;
    Trace tr;
    dev_free(dP_); dev_free(dQ_); dev_free(dPG_); dev_free(dQG_);
    dev_free(d_omega_p_); dev_free(d_omega_q_); dev_free(d_pmap_); dev_free(d_qmap_);
    dev_free(d_acc_); dev_free(d_err_); dev_free(d_outP_); dev_free(d_outQ_);
    dev_free(d_w0_); dev_free(d_w1_); dev_free(d_rr_); dev_free(d_goff_); dev_free(d_flags_); dev_free(d_tlock_);
    dev_free(d_R_); dev_free(d_order_); dev_free(d_e2_); dev_free(d_va_); dev_free(d_hidden_); dev_free(d_cv_raw_);
    tr.mark("release: device buffers");
    pinned_acc_give(h_acc_);
    h_acc_ = nullptr;
    pinned_give(h_order_pinned_);
    h_order_pinned_ = nullptr;
    pinned_give(h_neg_pinned_);
    h_neg_pinned_ = nullptr;
    dev_free(d_neg_);
    for (void *e : kernel_done_) cudaEventDestroy((cudaEvent_t)e);
    for (void *e : comm_done_) cudaEventDestroy((cudaEvent_t)e);
    kernel_done_.clear();
    comm_done_.clear();
    comm_ = nullptr;  // the communicator belongs to the process-wide cache (comm_for), not to the session
    if (comm_stream_) cudaStreamDestroy((cudaStream_t)comm_stream_);
    comm_stream_ = nullptr;
    if (ev0_) cudaEventDestroy((cudaEvent_t)ev0_);
    if (ev1_) cudaEventDestroy((cudaEvent_t)ev1_);
    tr.mark("release: pinned, events");
    if (stream_) {
        cudaStreamSynchronize((cudaStream_t)stream_);  // the frees above are ordered on this stream
        cudaStreamDestroy((cudaStream_t)stream_);
    }
    tr.mark("release: stream");
    t_pool_stream = nullptr;
    ev0_ = ev1_ = stream_ = nullptr;
    device_ready_ = loaded_ = false;
}

// One NCCL communicator per (world, rank, device) and process, created by the first session and kept: creating one
// costs seconds (ncclCommInitRank plus the first send/recv on every peer connection -- measured 2.2 s of "preprocessing"
// and 1.9 s for the first five epochs at 8 GPUs), a long-lived worker process trains many models.  All ranks of a job
// create their sessions in the same order, so either all of them reuse or all of them initialise; the unique id of a
// later session is ignored when its communicator already exists.
void *Session::comm_for(int world, int rank, int device, const unsigned char *id128) {
    struct Entry {
        int world, rank, device;
        void *comm;
    };
    static std::mutex mu;
    static std::vector<Entry> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Entry &e : cache)
            if (e.world == world && e.rank == rank && e.device == device) return e.comm;
    }
    const NcclApi *nc = nccl_api();
    if (!nc) return nullptr;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    // (outside the lock: ncclCommInitRank returns when every rank has joined, and the ranks may be threads of this process)
    ncclResult_t r = nc->CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error(std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(mu);
    cache.push_back(Entry{world, rank, device, (void *)comm});
    return (void *)comm;
}

int Session::init_device() {
    if (device_ready_) return 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        set_error("no CUDA device available: this build has no CPU fallback");
        return 1;
    }
    device_ = prm_.device >= 0 ? prm_.device : env_int("MFB200_DEVICE", -1);
    if (device_ < 0) CK(cudaGetDevice(&device_));
    CK(cudaSetDevice(device_));
    sm_count_ = mfk_sm_count(device_);
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    stream_ = st;
    t_pool_stream = st;
    if (pool_setup(device_)) return 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    ev0_ = e0;
    ev1_ = e1;
    if (dev_alloc(&d_acc_, kAccSize)) return 1;
    if (dev_alloc(&d_err_, 1)) return 1;
    if (pinned_acc_take(&h_acc_, kAccSize)) return 1;
    {  // the pinned staging buffers of this device (first use in the process: two cudaMallocHost of 32 MB).  Before the
       // communicator: once collectives are in flight no rank may sit in a call that synchronises devices.
        Staging &sg = staging();
        std::lock_guard<std::mutex> lock(sg.mu);
        sg.init();
    }
    if (world_ > 1) {  // one rank per GPU (processes, or threads of one process): the communicator of the S-stripe rotation
        const NcclApi *nc = nccl_api();
        if (!nc) return 1;
        comm_ = comm_for(world_, rank_, device_, nccl_id_);
        if (!comm_) return 1;
        cudaStream_t cs;
        CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        comm_stream_ = cs;
        for (int i = 0; i < 4 * world_; i++) {
            cudaEvent_t a, b;
            CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            kernel_done_.push_back(a);
            comm_done_.push_back(b);
        }
    }
    device_ready_ = true;
    return 0;
}

// gen_random_map (mf/mf.cpp:1009-1017): srand(0), identity, then the random_shuffle recurrence
// j = rand() % (i+1).  Uses the C library generator itself, as the reference does (and like the
// reference it therefore re-seeds the process-wide rand() state).
static std::vector<int> gen_map(int size) {
    std::srand(0);
    std::vector<int> a(size);
    for (int i = 0; i < size; i++) a[i] = i;
    for (int i = 1; i < size; i++) {
        const int j = std::rand() % (i + 1);
        if (i != j) std::swap(a[i], a[j]);
    }
    return a;
}

int Session::load(const mfb200_node *R, long long nnz) {
    t_pool_stream = (cudaStream_t)stream_;
    const double t0 = now_ms();
    if (world_ > 64) {  // one node; the sharded load keeps a rating's owner in a byte
        set_error("at most 64 ranks");
        return 1;
    }
    if (init_device()) return 1;
    CK(cudaSetDevice(device_));
    create_ms_ = now_ms() - t0;
    if (m_ < 0 || n_ < 0 || nnz < 0 || (nnz > 0 && !R)) {
        set_error("invalid problem");
        return 1;
    }
    if (nnz >= 0xffffffffll) {
        set_error("more than 2^32-1 ratings per device are not supported");
        return 1;
    }
    nnz_ = nnz;
    fun_ = prm_.fun;
    bpr_ = fun_ == MFK_FUN_ROW_BPR || fun_ == MFK_FUN_COL_BPR;
    if (fun_ != MFK_FUN_L2_MFR && fun_ != MFK_FUN_L1_MFR && fun_ != MFK_FUN_KL_MFR && fun_ != MFK_FUN_LR_MFC &&
        fun_ != MFK_FUN_L2_MFC && fun_ != MFK_FUN_L1_MFC && !bpr_) {
        set_error("loss function not supported (fun must be 0, 1, 2, 5, 6, 7, 10 or 11)");
        return 1;
    }
    if (prm_.lambda_p1 < 0 || prm_.lambda_q1 < 0) {
        set_error("regularization coefficient must be non-negative");
        return 1;
    }
    regression_ = fun_ == MFK_FUN_L2_MFR || fun_ == MFK_FUN_L1_MFR || fun_ == MFK_FUN_KL_MFR;
    mode_ = prm_.mode;
    if (mode_ == MFB200_MODE_AUTO)
        mode_ = nnz <= (long long)env_int("MFB200_EXACT_MAX_NNZ", 262144) ? MFB200_MODE_EXACT : MFB200_MODE_RING;
    if (bpr_) {
        // the one-class BPR losses (mf/mf.cpp:2131-2707) run in the exact mode only: every update touches a third row,
        // the negative the reference's scheduler draws, which the throughput schedule has no place for
        if (prm_.mode == MFB200_MODE_RING || prm_.mode == MFB200_MODE_RING_REPRO || world_ > 1 || !hidden_.empty() || va_nnz_ > 0) {
            set_error("the BPR losses run in the exact mode on one device, without validation set or cross-validation");
            return 1;
        }
        mode_ = MFB200_MODE_EXACT;
        draw_block_generators();  // (before the permutations re-seed rand(), as the reference's Scheduler constructor does)
    }
    reproducible_ = mode_ == MFB200_MODE_RING_REPRO || env_int("MFB200_REPRODUCIBLE", 0) != 0;
    if (mode_ == MFB200_MODE_RING_REPRO) mode_ = MFB200_MODE_RING;
    if (k_al_ > 512 && mode_ == MFB200_MODE_RING) mode_ = MFB200_MODE_EXACT;
    if (world_ > 1 && mode_ != MFB200_MODE_RING) {
        set_error("more than one GPU needs the throughput (ring) mode");
        return 1;
    }
    rowsP_alloc_ = (size_t)m_;
    rowsQ_alloc_ = (size_t)n_;
    if (mode_ == MFB200_MODE_RING) {
        // which throughput kernel: the run kernel (sgd_run.cu) or, for small launches, the cell kernel (sgd_cell.cu) for
        // the default loss at k_al <= 128, else the band kernel.  Cells need locks.
        const char *kn = std::getenv("MFB200_KERNEL");
        const bool supported = k_al_ <= 128 &&
                               mfk_sgd_run_supported(k_al_, 8, fun_, prm_.lambda_p1, prm_.lambda_q1, prm_.do_nmf ? 1 : 0) != 0;
        int kind = !supported ? 0 : reproducible_ ? 1 : 4;
        if (kn && !std::strcmp(kn, "band")) kind = 0;
        if (kn && !std::strcmp(kn, "run") && supported) kind = 1;
        if (kn && !std::strcmp(kn, "cell") && supported && !reproducible_) kind = 2;
        if (kn && !std::strcmp(kn, "warp") && supported && !reproducible_) kind = 3;
        if (kn && !std::strcmp(kn, "tlock") && supported && !reproducible_) kind = 5;
        if (kn && !std::strcmp(kn, "item") && supported && !reproducible_) kind = 6;
        if (!plan_band(m_, n_, nnz_, k_al_, sm_count_, mfk_sgd_band_max_smem(device_), world_, rank_, &plan_, kind)) return 1;
        // rows are padded so that every rank's T band and every S stripe has the same size (all-gather)
        const size_t rowsS = (size_t)plan_.nStripes * plan_.stripeRows, rowsT = (size_t)world_ * plan_.tSeg;
        rowsP_alloc_ = std::max(rowsP_alloc_, plan_.swap_sides ? rowsS : rowsT);
        rowsQ_alloc_ = std::max(rowsQ_alloc_, plan_.swap_sides ? rowsT : rowsS);
    }

    Trace tr;
    // The two permutations are a sequential walk over glibc's rand() (7.6 ms at the Netflix shape): in band mode they
    // are generated by a helper thread while the rating array travels to the device (upload_maps joins it).
    auto make_maps = [this] {
        if (shared_maps_) {  // ranks as threads of one process: rand() is walked by exactly one of them
            SharedMaps &sm = *shared_maps_;
            std::call_once(sm.once, [&] {
                sm.p = gen_map(m_);
                sm.q = gen_map(n_);
            });
            p_map_ = sm.p;
            q_map_ = sm.q;
        } else {
            p_map_ = gen_map(m_);
            q_map_ = gen_map(n_);
        }
    };
    if (mode_ == MFB200_MODE_RING) {
        map_thread_ = std::thread(make_maps);
    } else {
        make_maps();
        tr.mark("load: permutations (host)");
    }
    struct JoinGuard {  // no exit path may leave the helper thread running
        std::thread &t;
        ~JoinGuard() {
            if (t.joinable()) t.join();
        }
    } join_guard{map_thread_};
    if (dev_alloc(&d_pmap_, (size_t)m_) || dev_alloc(&d_qmap_, (size_t)n_)) return 1;
    if (mode_ != MFB200_MODE_RING && upload_maps()) return 1;
    if (dev_alloc(&dP_, rowsP_alloc_ * k_al_) || dev_alloc(&dQ_, rowsQ_alloc_ * k_al_) ||
        dev_alloc(&dPG_, rowsP_alloc_ * 2) || dev_alloc(&dQG_, rowsQ_alloc_ * 2) ||
        dev_alloc(&d_omega_p_, (size_t)m_) || dev_alloc(&d_omega_q_, (size_t)n_))
        return 1;
    CK(cudaMemsetAsync(d_omega_p_, 0, sizeof(int) * (size_t)std::max(m_, 1), (cudaStream_t)stream_));
    CK(cudaMemsetAsync(d_omega_q_, 0, sizeof(int) * (size_t)std::max(n_, 1), (cudaStream_t)stream_));

    int rc = mode_ == MFB200_MODE_EXACT ? load_exact(R) : load_band(R);
    if (rc) return rc;
    // lambda rescaling of fpsg_core, mf/mf.cpp:2798-2816 (float divisions)
    lambda_p_ = prm_.lambda_p2;
    lambda_q_ = prm_.lambda_q2;
    lambda_p1_ = prm_.lambda_p1;
    lambda_q1_ = prm_.lambda_q1;
    if (fun_ == MFK_FUN_L2_MFR) {
        lambda_p_ /= scale_;
        lambda_q_ /= scale_;
        lambda_p1_ /= (float)std::pow(scale_, 1.5);
        lambda_q1_ /= (float)std::pow(scale_, 1.5);
    } else if (fun_ == MFK_FUN_L1_MFR || fun_ == MFK_FUN_KL_MFR) {
        lambda_p1_ /= std::sqrt(scale_);
        lambda_q1_ /= std::sqrt(scale_);
    }
    tr.mark("load: mode preprocessing");
    if (init_model()) return 1;
    if (va_nnz_ > 0) {
        if (dev_alloc(&d_va_, (size_t)va_nnz_)) return 1;
        CK(cudaMemcpyAsync(d_va_, va_host_, sizeof(mfk_node) * (size_t)va_nnz_, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    }
    CK(cudaStreamSynchronize((cudaStream_t)stream_));
    tr.mark("load: init model");
    loaded_ = true;
    prep_ms_ = now_ms() - t0;
    return 0;
}

int Session::upload_maps() {
    if (map_thread_.joinable()) map_thread_.join();
    CK(cudaMemcpyAsync(d_pmap_, p_map_.data(), sizeof(int) * (size_t)m_, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    CK(cudaMemcpyAsync(d_qmap_, q_map_.data(), sizeof(int) * (size_t)n_, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    return 0;
}

// ---- exact mode: reproduce steps 1-5 of SURVEY.md Appendix A.2 on the host (ordering only) -------
int Session::load_exact(const mfb200_node *R) {
    const int bins = std::max(1, prm_.nr_bins), nblk = bins * bins;
    hR_.resize((size_t)nnz_);
    // collect_info, mf/mf.cpp:462-484 (sequential double sums, as at nr_threads = 1)
    double ex = 0, ex2 = 0;
    for (long long i = 0; i < nnz_; i++) {
        ex += (double)R[i].r;
        ex2 += (double)R[i].r * R[i].r;
    }
    if (nnz_ > 0) {
        ex /= (double)nnz_;
        ex2 /= (double)nnz_;
    }
    avg_ = (float)ex;
    std_dev_ = nnz_ > 0 ? (float)std::sqrt(ex2 - ex * ex) : 0.f;
    scale_ = regression_ ? std::max(1e-4f, std_dev_) : 1.0f;  // mf/mf.cpp:2996-2999: only the regression losses scale
    const float inv = 1.0f / scale_;     // mf/mf.cpp:3010
    for (long long i = 0; i < nnz_; i++) {
        if (R[i].u < 0 || R[i].u >= m_ || R[i].v < 0 || R[i].v >= n_) {
            set_error("rating index out of range");
            return 1;
        }
        hR_[i].u = p_map_[R[i].u];
        hR_[i].v = q_map_[R[i].v];
        hR_[i].r = inv == 1.0f ? R[i].r : R[i].r * inv;
    }
    // grid_problem, mf/mf.cpp:793-858
    const int seg_p = std::max(1, (int)std::ceil((double)m_ / bins)), seg_q = std::max(1, (int)std::ceil((double)n_ / bins));
    auto home = [&](const mfk_node &N) { return (N.u / seg_p) * bins + N.v / seg_q; };
    std::vector<long long> cnt(nblk, 0);
    std::vector<int> omega_p(std::max(m_, 1), 0), omega_q(std::max(n_, 1), 0);
    for (long long i = 0; i < nnz_; i++) {
        cnt[home(hR_[i])]++;
        omega_p[hR_[i].u]++;
        omega_q[hR_[i].v]++;
    }
    blk_first_.assign(nblk + 1, 0);
    for (int b = 0; b < nblk; b++) blk_first_[b + 1] = blk_first_[b] + cnt[b];
    {
        std::vector<long long> fill(blk_first_.begin(), blk_first_.end() - 1);
        for (int b = 0; b < nblk; b++)
            for (long long at = fill[b]; at != blk_first_[b + 1];) {
                const int h = home(hR_[at]);
                if (h == b)
                    at++;
                else
                    std::swap(hR_[at], hR_[fill[h]++]);
            }
    }
    const bool by_u = m_ > n_;
    for (int b = 0; b < nblk; b++)
        std::sort(hR_.begin() + blk_first_[b], hR_.begin() + blk_first_[b + 1],
                  [by_u](const mfk_node &x, const mfk_node &y) {
                      if (by_u) return x.u != y.u ? x.u < y.u : x.v < y.v;
                      return x.v != y.v ? x.v < y.v : x.u < y.u;
                  });

    cudaStream_t st = (cudaStream_t)stream_;
    if (dev_alloc(&d_R_, (size_t)nnz_) || dev_alloc(&d_order_, (size_t)nnz_) || dev_alloc(&d_e2_, 2 * (size_t)nnz_)) return 1;  // loss terms, then the hinge losses' correct-sign flags
    CK(cudaMemsetAsync(d_e2_, 0, sizeof(float) * 2 * (size_t)nnz_, st));
    if (pinned_take((void **)&h_order_pinned_, sizeof(unsigned) * (size_t)std::max<long long>(nnz_, 1))) return 1;
    if (bpr_) {
        if (pinned_take((void **)&h_neg_pinned_, sizeof(int) * (size_t)std::max<long long>(nnz_, 1))) return 1;
        if (dev_alloc(&d_neg_, (size_t)nnz_)) return 1;
    }
    CK(cudaMemcpyAsync(d_R_, hR_.data(), sizeof(mfk_node) * (size_t)nnz_, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_omega_p_, omega_p.data(), sizeof(int) * (size_t)m_, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_omega_q_, omega_q.data(), sizeof(int) * (size_t)n_, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // omega vectors go out of scope
    lvl_u_.assign(std::max(m_, 1), 0);
    lvl_v_.assign(std::max(n_, 1), 0);
    return 0;
}

// ---- band mode: steps 1-5 on the device -----------------------------------------------------------
int Session::load_band(const mfb200_node *R) {
    cudaStream_t st = (cudaStream_t)stream_;
    const mfk_band_shape &sh = plan_;
    Trace tr;
    mfk_node *d_raw = nullptr;
    unsigned long long *d_k0 = nullptr, *d_k1 = nullptr, *d_x0 = nullptr, *d_x1 = nullptr, *d_kept = nullptr;
    unsigned *d_v0 = nullptr, *d_v1 = nullptr, *d_first = nullptr;
    int *d_bad = nullptr;
    void *d_tmp = nullptr;
    const size_t n_off = (size_t)sh.nStripes * sh.nC * sh.nPass * (sh.by_row == 2 ? sh.nTB : sh.nG) + 1;
    // Several GPUs: every rank is handed the whole rating array but uploads only its 1/world slice, finds the rank that
    // owns each rating's T row and ships the ratings there (ncclSend/ncclRecv); statistics and omega are summed over
    // the ranks.  Measured with every rank uploading everything: 211 ms of preprocessing at 4 GPUs against 37-60 at one.
    const bool sharded = world_ > 1 && hidden_.empty() && env_int("MFB200_SHARDED_LOAD", 1) != 0;
    long long cnt_in = nnz_;  // ratings this rank's key kernels look at
    mfk_node *d_slice = nullptr, *d_grouped = nullptr;
    unsigned char *d_own0 = nullptr, *d_own1 = nullptr;
    unsigned long long *d_cnt = nullptr, *d_cnt_all = nullptr;
    void *d_gtmp = nullptr;
    int rc = 1;
    do {
        const long long s_lo = sharded ? nnz_ * rank_ / world_ : 0, s_hi = sharded ? nnz_ * (rank_ + 1) / world_ : nnz_;
        const long long ns = s_hi - s_lo;
        mfk_node *&d_first_copy = sharded ? d_slice : d_raw;
        if (dev_alloc(&d_first_copy, (size_t)ns)) break;
        if (staged_h2d(d_first_copy, R + s_lo, sizeof(mfk_node) * (size_t)ns, st)) break;
        tr.mark("band: H2D ratings");
        // collect_info on the device (double sums)
        if (cudaMemsetAsync(d_acc_, 0, sizeof(double) * 8, st) != cudaSuccess) break;
        if (mfk_stats(d_first_copy, ns, d_acc_, st)) break;
        if (sharded && nccl_api()->AllReduce(d_acc_, d_acc_, 2, ncclFloat64, ncclSum, (ncclComm_t)comm_, st) != ncclSuccess) {
            set_error("ncclAllReduce (rating statistics) failed");
            rc = 2;
            break;
        }
        if (cudaMemcpyAsync(h_acc_, d_acc_, sizeof(double) * 2, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        double ex = h_acc_[0], ex2 = h_acc_[1];
        if (nnz_ > 0) {
            ex /= (double)nnz_;
            ex2 /= (double)nnz_;
        }
        avg_ = (float)ex;
        std_dev_ = nnz_ > 0 ? (float)std::sqrt(std::max(0.0, ex2 - ex * ex)) : 0.f;
        scale_ = regression_ ? std::max(1e-4f, std_dev_) : 1.0f;  // mf/mf.cpp:2996-2999
        const float inv = 1.0f / scale_;
        tr.mark("band: stats");
        if (upload_maps()) break;
        if (dev_alloc(&d_kept, 1) || dev_alloc(&d_bad, 1)) break;
        if (cudaMemsetAsync(d_bad, 0, sizeof(int), st) != cudaSuccess) break;

        if (sharded) {
            const NcclApi *nc = nccl_api();
            const int W = world_, me = rank_;
            if (dev_alloc(&d_own0, (size_t)ns) || dev_alloc(&d_own1, (size_t)ns) || dev_alloc(&d_grouped, (size_t)ns) ||
                dev_alloc(&d_cnt, (size_t)W + 2) || dev_alloc(&d_cnt_all, (size_t)W * (W + 2)))
                break;
            if (cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * (size_t)(W + 2), st) != cudaSuccess) break;
            if (mfk_owner_of(d_slice, ns, d_pmap_, d_qmap_, sh.swap_sides, sh.tSeg, W, d_omega_p_, d_omega_q_, d_own0, d_cnt,
                             d_bad, m_, n_, st))
                break;
            // counts[W + 1] = this rank's out-of-range flag, so that all ranks fail together
            if (cudaMemcpyAsync(d_cnt + W + 1, d_bad, sizeof(int), cudaMemcpyDeviceToDevice, st) != cudaSuccess) break;
            bool ok = nc->GroupStart() == ncclSuccess;
            ok = ok && nc->AllReduce(d_omega_p_, d_omega_p_, (size_t)m_, ncclInt32, ncclSum, (ncclComm_t)comm_, st) == ncclSuccess;
            ok = ok && nc->AllReduce(d_omega_q_, d_omega_q_, (size_t)n_, ncclInt32, ncclSum, (ncclComm_t)comm_, st) == ncclSuccess;
            ok = ok && nc->AllGather(d_cnt, d_cnt_all, (size_t)W + 2, ncclUint64, (ncclComm_t)comm_, st) == ncclSuccess;
            ok = ok && nc->GroupEnd() == ncclSuccess;
            if (!ok) {
                set_error("NCCL exchange of the rating counts failed");
                rc = 2;
                break;
            }
            std::vector<unsigned long long> cnt((size_t)W * (W + 2));
            if (cudaMemcpyAsync(cnt.data(), d_cnt_all, sizeof(unsigned long long) * cnt.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
            if (cudaStreamSynchronize(st) != cudaSuccess) break;
            bool any_bad = false;
            for (int q = 0; q < W; q++) any_bad = any_bad || cnt[(size_t)q * (W + 2) + W + 1] != 0;
            if (any_bad) {
                set_error("rating index out of range");
                rc = 2;
                break;
            }
            // group the slice by destination, then the all-to-all: rank q's block for me lands at the sum of the blocks of
            // the ranks before q
            const size_t gbytes = mfk_group_tmp_bytes(ns);
            if (cudaMallocAsync(&d_gtmp, gbytes ? gbytes : 1, st) != cudaSuccess) break;
            if (mfk_group_by_owner(d_own0, d_own1, d_slice, d_grouped, ns, bits_for(W + 1), d_gtmp, gbytes, st)) break;
            std::vector<unsigned long long> cmat((size_t)W * W);  // counts[q][d] without the two extra columns
            for (int q = 0; q < W; q++)
                for (int d = 0; d < W; d++) cmat[(size_t)q * W + d] = cnt[(size_t)q * (W + 2) + d];
            std::vector<long long> soff((size_t)W), roff((size_t)W);
            const long long nr = exchange_plan(W, me, cmat.data(), soff.data(), roff.data());
            if (dev_alloc(&d_raw, (size_t)nr)) break;
            ok = nc->GroupStart() == ncclSuccess;
            for (int q = 0; q < W && ok; q++) {
                const long long sc = (long long)cmat[(size_t)me * W + q], rcq = (long long)cmat[(size_t)q * W + me];
                if (q == me) {
                    ok = cudaMemcpyAsync(d_raw + roff[q], d_grouped + soff[q], sizeof(mfk_node) * (size_t)sc, cudaMemcpyDeviceToDevice, st) == cudaSuccess;
                } else {
                    if (sc > 0) ok = ok && nc->Send(d_grouped + soff[q], sizeof(mfk_node) * (size_t)sc, ncclChar, q, (ncclComm_t)comm_, st) == ncclSuccess;
                    if (rcq > 0) ok = ok && nc->Recv(d_raw + roff[q], sizeof(mfk_node) * (size_t)rcq, ncclChar, q, (ncclComm_t)comm_, st) == ncclSuccess;
                }
            }
            ok = ok && nc->GroupEnd() == ncclSuccess;
            if (!ok) {
                set_error("NCCL exchange of the ratings failed");
                rc = 2;
                break;
            }
            cnt_in = nr;
            tr.mark("band: ratings to their owners");
        }

        if (dev_alloc(&d_k0, (size_t)cnt_in) || dev_alloc(&d_k1, (size_t)cnt_in) || dev_alloc(&d_v0, (size_t)cnt_in) ||
            dev_alloc(&d_v1, (size_t)cnt_in) || dev_alloc(&d_x0, (size_t)cnt_in) || dev_alloc(&d_x1, (size_t)cnt_in) ||
            dev_alloc(&d_first, (size_t)std::min(m_, n_) + 1))
            break;
        if (dev_alloc(&d_goff_, n_off) || dev_alloc(&d_flags_, (size_t)sh.nC * sh.nG)) break;
        if (sh.tlock && dev_alloc(&d_tlock_, (size_t)std::max(1, sh.tSeg))) break;
        if (cudaMemsetAsync(d_goff_, 0, sizeof(unsigned) * n_off, st) != cudaSuccess) break;
        if (cudaMemsetAsync(d_kept, 0, sizeof(unsigned long long), st) != cudaSuccess) break;
        tr.mark("band: alloc work buffers");
        nnz_kept_ = 0;
        if (cnt_in > 0) {
            if (upload_hidden_mask()) break;
            // (sharded: omega has been counted on the slices and summed; the received ratings all belong to this rank)
            if (mfk_band_keys1(d_raw, cnt_in, d_pmap_, d_qmap_, sh, inv, sharded ? nullptr : d_omega_p_,
                               sharded ? nullptr : d_omega_q_, d_k0, d_x0, d_kept, d_bad, m_, n_, hidden_arg(), st))
                break;
            unsigned long long kept = 0;
            int bad = 0;
            if (cudaMemcpyAsync(&kept, d_kept, sizeof(kept), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
            if (cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
            if (cudaStreamSynchronize(st) != cudaSuccess) break;
            if (bad) {
                set_error("rating index out of range");
                rc = 2;
                break;
            }
            nnz_kept_ = (long long)kept;
            if (hidden_.empty()) {
                dev_free(d_raw);  // everything needed is in the keys and payloads now
            } else {
                d_cv_raw_ = d_raw;  // cross-validation: the hidden ratings are evaluated after training (cv_error)
                d_raw = nullptr;
            }
            tr.mark("band: keys (stream order) + omega");
            const size_t tmp_bytes = mfk_sort_tmp_bytes(cnt_in);
            if (cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 1, st) != cudaSuccess) break;
            const int low = sh.bitsT + sh.bitsD + sh.bitsA;
            if (mfk_sort_pairs64(d_k0, d_k1, d_x0, d_x1, cnt_in, std::min(64, sh.bitsSB + sh.bitsG + low + 1), d_tmp, tmp_bytes, st))
                break;
            tr.mark("band: radix sort 1");
            // d_k1/d_x1: the stream.  Tickets (the per-row update order) are only needed by the reproducible
            // variant; with locks the second sort is skipped and the ticket field of the stream stays 0.
            if (reproducible_) {
                // ranks -> ticket-order keys in d_k0, stream positions in d_v0
                if (mfk_band_segstart(d_k1, nnz_kept_, sh, d_v0, d_v1, d_tmp, tmp_bytes, st)) break;
                if (mfk_band_keys2(d_k1, d_x1, d_v1, nnz_kept_, sh, d_k0, d_v0, st)) break;
                tr.mark("band: ranks + keys (ticket order)");
                const int bits2 = sh.bitsB + sh.bitsT + mfk_band_rank_bits(sh) + sh.bitsG;
                if (mfk_sort_pairs32(d_k0, (unsigned long long *)d_x0, d_v0, d_v1, nnz_kept_, std::min(64, bits2), d_tmp, tmp_bytes, st))
                    break;
                tr.mark("band: radix sort 2");
                if (mfk_band_tickets((unsigned long long *)d_x0, d_v1, nnz_kept_, sh, d_first, d_v0, st)) break;
            }
            if (dev_alloc(&d_w0_, (size_t)nnz_kept_) || dev_alloc(&d_w1_, (size_t)nnz_kept_) || dev_alloc(&d_rr_, (size_t)nnz_kept_))
                break;
            if (mfk_band_stream(d_k1, d_x1, reproducible_ ? d_v0 : nullptr, nnz_kept_, sh, d_w0_, d_w1_, d_rr_, d_goff_, st)) break;
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        tr.mark("band: stream + offsets");
        rc = 0;
    } while (0);
    if (rc == 1) {
        cudaError_t e = cudaGetLastError();
        set_error(std::string("band preprocessing failed: ") + cudaGetErrorString(e));
    }
    dev_free(d_raw); dev_free(d_k0); dev_free(d_k1); dev_free(d_v0); dev_free(d_v1); dev_free(d_x0); dev_free(d_x1);
    dev_free(d_first); dev_free(d_kept); dev_free(d_bad);
    dev_free(d_slice); dev_free(d_grouped); dev_free(d_own0); dev_free(d_own1); dev_free(d_cnt); dev_free(d_cnt_all);
    if (d_tmp) cudaFreeAsync(d_tmp, st);
    if (d_gtmp) cudaFreeAsync(d_gtmp, st);
    return rc ? 1 : 0;
}

// init_model (mf/mf.cpp:952-1007) + PG/QG = 1 (2835) + scheduler state.
int Session::init_model() {
    cudaStream_t st = (cudaStream_t)stream_;
    int *d_rank = nullptr, *d_total = nullptr;
    void *d_tmp = nullptr;
    const int rows_max = std::max(std::max(m_, n_), 1);
    const size_t tmp_bytes = mfk_rank_tmp_bytes(rows_max);
    if (dev_alloc(&d_rank, (size_t)rows_max) || dev_alloc(&d_total, 1)) return 1;
    CK(cudaMallocAsync(&d_tmp, tmp_bytes, st));
    int total_p = 0;
    CK(mfk_exclusive_rank(d_omega_p_, m_, d_rank, d_total, d_tmp, tmp_bytes, st));
    CK(mfk_init_rows(dP_, dPG_, d_omega_p_, d_rank, 0, m_, k_, k_al_, bpr_ ? 1 : 0, st));
    CK(cudaMemcpyAsync(&total_p, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(mfk_exclusive_rank(d_omega_q_, n_, d_rank, d_total, d_tmp, tmp_bytes, st));
    CK(mfk_init_rows(dQ_, dQG_, d_omega_q_, d_rank, total_p, n_, k_, k_al_, bpr_ ? 1 : 0, st));
    CK(cudaStreamSynchronize(st));
    cudaFreeAsync(d_tmp, st);
    dev_free(d_rank);
    dev_free(d_total);

    epochs_done_ = 0;
    launches_ = 0;
    epochs_ms_ = 0;
    header_printed_ = false;
    CK(cudaMemsetAsync(d_err_, 0, sizeof(int), st));
    if (mode_ == MFB200_MODE_RING) {
        CK(cudaMemsetAsync(d_flags_, 0, sizeof(unsigned) * (size_t)plan_.nC * plan_.nG, st));
        if (d_tlock_) CK(cudaMemsetAsync(d_tlock_, 0, sizeof(unsigned) * (size_t)std::max(1, plan_.tSeg), st));
        step_base_ = 0;
    } else {
        // Scheduler constructor (mf/mf.cpp:89-111): its own default-seeded engine draws one priority
        // per block; a min-heap on (priority, block id).
        const int nblk = std::max(1, prm_.nr_bins) * std::max(1, prm_.nr_bins);
        sched_rng_ = std::default_random_engine();
        std::uniform_real_distribution<float> dist(0.0f, 1.0f);
        heap_ = decltype(heap_)();
        for (int i = 0; i < nblk; i++)
            if (!is_hidden(i)) heap_.emplace(dist(sched_rng_), i);  // hidden blocks get no priority and no draw (104-111)
        visits_.assign(nblk, 0);
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

// Scheduler's constructor (mf/mf.cpp:103-110) seeds one minstd_rand0 per grid block from the PROCESS-WIDE rand(), whatever
// state the host program left it in -- the same calls to the same C library here, so a host that seeds rand() gets the
// negatives the reference would give it.  minstd_rand0(s): state s mod (2^31 - 1), 1 if that is 0.
void Session::draw_block_generators() {
    const int nblk = std::max(1, prm_.nr_bins) * std::max(1, prm_.nr_bins);
    block_gen_.assign((size_t)nblk, 1u);
    for (int i = 0; i < nblk; i++) {
        const unsigned sd = (unsigned)std::rand() % 2147483647u;
        block_gen_[(size_t)i] = sd ? sd : 1u;
    }
}

// Scheduler::get_negative (mf/mf.cpp:249-280): one draw of the first block's generator; odd -> a row of the first block's
// range, even -> of the second block's
int Session::bpr_negative(int first_block, int second_block) {
    const int bins = std::max(1, prm_.nr_bins);
    const bool col = fun_ == MFK_FUN_COL_BPR;
    unsigned &g = block_gen_[(size_t)first_block];
    g = (unsigned)(((unsigned long long)g * 16807ull) % 2147483647ull);
    const int rand_val = (int)g;
    auto gen_random = [&](int block_id) {
        int v_min, v_max;
        if (col) {
            const int seg = (int)std::ceil((double)m_ / bins);
            v_min = std::min((block_id / bins) * seg, m_ - 1);
            v_max = std::min(v_min + seg, m_ - 1);
        } else {
            const int seg = (int)std::ceil((double)n_ / bins);
            v_min = std::min((block_id % bins) * seg, n_ - 1);
            v_max = std::min(v_min + seg, n_ - 1);
        }
        return v_max == v_min ? v_min : rand_val % (v_max - v_min) + v_min;
    };
    return (rand_val % 2) ? gen_random(first_block) : gen_random(second_block);
}

int Session::reset() {
    t_pool_stream = (cudaStream_t)stream_;
    if (!loaded_) {
        set_error("session not loaded");
        return 1;
    }
    CK(cudaSetDevice(device_));
    if (bpr_) draw_block_generators();
    return init_model();
}

// One epoch of the reference's single-thread loop = bins^2 block jobs in scheduler order
// (mf/mf.cpp:113-150 pop, 193-207 push), every block's ratings in stored order (1220-1235).
// level(rating) = 1 + max(level of the previous rating with the same row, same column): ratings of
// one level are independent, so they run as one launch; levels run in order.
int Session::epoch_exact(double *loss_out, double *err_out) {
    cudaStream_t st = (cudaStream_t)stream_;
    const int nblk = (int)visits_.size();
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    // mf/mf.cpp:2834, 2910-2911: dims 0-7 only in the first epoch, unless an L1 term is on
    const int slow_only = (epochs_done_ == 0 && lambda_p1_ == 0 && lambda_q1_ == 0) ? 1 : 0;
    const bool hinge = fun_ == MFK_FUN_L2_MFC || fun_ == MFK_FUN_L1_MFC;  // error != loss: count of correct signs
    // Visits are gathered into portions of at most nnz_ ratings (the size of the order buffers): without hidden blocks
    // an epoch is exactly one portion; with them (cross-validation) the nr_bins^2 jobs of an epoch go to fewer blocks,
    // some blocks are visited twice and the epoch may need a second portion.  Portions run one after the other on the
    // stream, so levels only have to be consistent inside a portion.
    std::vector<long long> seq;  // rating indices in processing order
    std::vector<int> level;      // level of every visit
    std::vector<int> negs;       // BPR: the negative row of every visit
    seq.reserve((size_t)nnz_);
    level.reserve((size_t)nnz_);
    if (bpr_) negs.reserve((size_t)nnz_);
    const bool col = fun_ == MFK_FUN_COL_BPR;
    int max_level = 0;
    auto flush = [&]() -> int {
        if (seq.empty()) return 0;
        std::vector<long long> first(max_level + 2, 0);
        for (size_t s = 0; s < seq.size(); s++) first[level[s] + 1]++;
        for (int l = 1; l <= max_level + 1; l++) first[l] += first[l - 1];
        {
            std::vector<long long> fill(first.begin(), first.end());
            for (size_t s = 0; s < seq.size(); s++) {
                const long long at = fill[level[s]]++;
                h_order_pinned_[at] = (unsigned)seq[s];
                if (bpr_) h_neg_pinned_[at] = negs[s];
            }
        }
        CK(cudaMemcpyAsync(d_order_, h_order_pinned_, sizeof(unsigned) * seq.size(), cudaMemcpyHostToDevice, st));
        if (bpr_) CK(cudaMemcpyAsync(d_neg_, h_neg_pinned_, sizeof(int) * seq.size(), cudaMemcpyHostToDevice, st));
        for (int l = 1; l <= max_level; l++) {
            const long long cntl = first[l + 1] - first[l];
            if (bpr_)  // COL_BPR_MFOC::load_fixed_variables swaps the coefficients with the rows (mf/mf.cpp:2645-2686)
                CK(mfk_bpr_exact_level(d_R_, d_order_, d_neg_, (int)first[l], (int)cntl, dP_, dQ_, dPG_, dQG_, k_al_,
                                       col ? lambda_q_ : lambda_p_, col ? lambda_p_ : lambda_q_, prm_.eta, slow_only, d_e2_,
                                       col ? 1 : 0, col ? lambda_q1_ : lambda_p1_, col ? lambda_p1_ : lambda_q1_,
                                       prm_.do_nmf, st));
            else
                CK(mfk_sgd_exact_level(d_R_, d_order_ + first[l], (int)cntl, dP_, dQ_, dPG_, dQG_, k_al_, lambda_p_, lambda_q_,
                                       prm_.eta, slow_only, d_e2_, fun_, lambda_p1_, lambda_q1_, prm_.do_nmf,
                                       hinge ? d_e2_ + nnz_ : nullptr, st));
            launches_++;
        }
        CK(cudaStreamSynchronize(st));  // the pinned order buffer is reused by the next portion
        seq.clear();
        level.clear();
        negs.clear();
        max_level = 0;
        std::fill(lvl_u_.begin(), lvl_u_.end(), 0);
        std::fill(lvl_v_.begin(), lvl_v_.end(), 0);
        return 0;
    };
    std::fill(lvl_u_.begin(), lvl_u_.end(), 0);
    std::fill(lvl_v_.begin(), lvl_v_.end(), 0);
    for (int job = 0; job < nblk; job++) {
        const Job top = heap_.top();
        heap_.pop();
        const int blk = top.second;
        visits_[blk]++;
        if ((long long)seq.size() + (blk_first_[blk + 1] - blk_first_[blk]) > nnz_ && flush()) return 1;
        if (bpr_) {
            // BPRSolver::arrange_block (mf/mf.cpp:2193-2201): a second block is taken out of the queue for the negatives --
            // Scheduler::get_bpr_job (152-191) at one thread: the first block in priority order that shares blk's row band
            // (column band when column-oriented) and not its other band
            const int bins = std::max(1, prm_.nr_bins);
            int second = blk;
            {
                std::vector<Job> locked;
                while (!heap_.empty()) {
                    const Job cand = heap_.top();
                    heap_.pop();
                    const int pb = cand.second / bins, qb = cand.second % bins;
                    const bool rejected = col ? (blk % bins != qb || pb == blk / bins) : (blk / bins != pb || qb == blk % bins);
                    if (rejected) {
                        locked.push_back(cand);
                    } else {
                        second = cand.second;
                        break;
                    }
                }
                for (const Job &j : locked) heap_.push(j);
            }
            for (long long i = blk_first_[blk]; i < blk_first_[blk + 1]; i++) {
                const int u = hR_[i].u, v = hR_[i].v;
                const int w = bpr_negative(blk, second);
                int &lw = col ? lvl_u_[w] : lvl_v_[w];  // the third row: another item (another user when column-oriented)
                const int l = std::max(std::max(lvl_u_[u], lvl_v_[v]), lw) + 1;
                lvl_u_[u] = lvl_v_[v] = l;
                lw = l;
                level.push_back(l);
                negs.push_back(w);
                if (l > max_level) max_level = l;
                seq.push_back(i);
            }
            heap_.emplace((float)visits_[blk] + dist(sched_rng_), blk);                              // put_job, 202-204
            if (second != blk) heap_.emplace((float)visits_[second] + dist(sched_rng_), second);    // put_bpr_job, 222-235
            continue;
        }
        for (long long i = blk_first_[blk]; i < blk_first_[blk + 1]; i++) {
            const int u = hR_[i].u, v = hR_[i].v;
            const int l = std::max(lvl_u_[u], lvl_v_[v]) + 1;
            lvl_u_[u] = lvl_v_[v] = l;
            level.push_back(l);
            if (l > max_level) max_level = l;
            seq.push_back(i);
        }
        heap_.emplace((float)visits_[blk] + dist(sched_rng_), blk);
    }
    if (flush()) return 1;
    // the loss terms are stored per rating: a block visited twice keeps the terms of its last visit, a hidden block
    // keeps zeros -- block_losses of the reference's scheduler (mf/mf.cpp:199-200, 237-247)
    CK(cudaMemsetAsync(d_acc_, 0, sizeof(double) * 2, st));
    CK(mfk_sum_f32(d_e2_, nnz_, d_acc_, st));
    if (hinge) CK(mfk_sum_f32(d_e2_ + nnz_, nnz_, d_acc_ + 1, st));
    CK(cudaMemcpyAsync(h_acc_, d_acc_, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *loss_out = h_acc_[0];
    *err_out = hinge ? h_acc_[1] : h_acc_[0];
    return 0;
}

int Session::epochs_band(int epochs, double *loss_out, double *err_out) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (epochs > 1024) {
        set_error("at most 1024 epochs per call");
        return 1;
    }
    CK(cudaMemsetAsync(d_acc_, 0, sizeof(double) * (size_t)epochs, st));
    CK(cudaMemsetAsync(d_acc_ + kAccErr, 0, sizeof(double) * (size_t)epochs, st));
    mfk_band_args a;
    std::memset(&a, 0, sizeof(a));
    const bool sw = plan_.swap_sides != 0;  // S = users when m < n
    a.S = sw ? dP_ : dQ_;
    a.T = sw ? dQ_ : dP_;
    a.SG = sw ? dPG_ : dQG_;
    a.TG = sw ? dQG_ : dPG_;
    a.lambda_s = sw ? lambda_p_ : lambda_q_;
    a.lambda_t = sw ? lambda_q_ : lambda_p_;
    a.eta = prm_.eta;
    a.fun = fun_;
    a.lambda1_s = sw ? lambda_p1_ : lambda_q1_;
    a.lambda1_t = sw ? lambda_q1_ : lambda_p1_;
    a.do_nmf = prm_.do_nmf ? 1 : 0;
    a.w0 = d_w0_;
    a.w1 = d_w1_;
    a.rr = d_rr_;
    a.goff = d_goff_;
    a.flags = d_flags_;
    a.tlock = plan_.tlock ? d_tlock_ : nullptr;
    a.error_flag = d_err_;
    unsigned long long *d_stats = nullptr;
    if (env_int("MFB200_STATS", 0)) {
        if (dev_alloc(&d_stats, 16)) return 1;
        CK(cudaMemsetAsync(d_stats, 0, sizeof(unsigned long long) * 16, st));
    }
    a.stats = d_stats;
    a.dynamic = reproducible_ ? 0 : 1;
    // few S rows per CTA (item stripes rotating over several GPUs, small problems): the hold time of a row bounds the
    // launch, so the lock is taken late and released early (kernels.cu, LATE); measured break-even ~1.5 rows per group
    a.late_lock = env_int("MFB200_LATE_LOCK", plan_.segS * 4 <= plan_.nG * 5 ? 1 : 0);
    a.shape = plan_;
    a.k_al = k_al_;
    // a hand-off that does not arrive within this many seconds of wall clock makes the launch give up (a lost
    // hand-off would be a bug; a stalled neighbour -- time-slicing, a hung peer GPU -- is not)
    a.wait_limit_ns = (unsigned long long)std::max(1, env_int("MFB200_WAIT_LIMIT_S", 30)) * 1000000000ull;
    const int nS_total = sw ? m_ : n_;
    const size_t n_off_stripe = (size_t)plan_.nC * plan_.nPass * (plan_.by_row == 2 ? plan_.nTB : plan_.nG);
    float *const S0 = a.S, *const SG0 = a.SG;
    a.T += (size_t)plan_.tLo * k_al_;  // the stream addresses T rows relative to this rank's band
    a.TG += (size_t)plan_.tLo * 2;
    const NcclApi *nc = world_ > 1 ? nccl_api() : nullptr;
    cudaStream_t cs = (cudaStream_t)comm_stream_;
    const int nsub = plan_.nStripes;  // launches per epoch: 1, or 2*world half-stripes
    for (int e = 0; e < epochs; e++) {
        a.full = ((epochs_done_ + e) > 0 || lambda_p1_ != 0 || lambda_q1_ != 0) ? 1 : 0;  // mf/mf.cpp:2834, 2910-2911
        a.loss = d_acc_ + e;
        a.err = d_acc_ + kAccErr + e;
        for (int sub = 0; sub < nsub; sub++) {
            int js = 0;
            RotationStep rs{};
            if (world_ > 1) {
                const int spr = nsub / world_;
                rs = rotation_step(world_, rank_, substeps_done_, spr);
                js = rs.compute;
                // the stripe trained now arrived with the transfer issued `spr` sub-steps ago
                if (substeps_done_ >= spr) CK(cudaStreamWaitEvent(st, (cudaEvent_t)comm_done_[(substeps_done_ - spr) % nsub], 0));
            }
            const int row0 = js * plan_.stripeRows;
            a.S = S0 + (size_t)row0 * k_al_;
            a.SG = SG0 + (size_t)row0 * 2;
            a.nS = std::max(0, std::min(plan_.stripeRows, nS_total - row0));
            a.goff = d_goff_ + (size_t)js * n_off_stripe;
            a.base = step_base_;
            if (nnz_kept_ > 0 && a.nS > 0)
                CK(plan_.by_row == 4 ? mfk_sgd_item_epoch(&a, st) : plan_.by_row == 3 ? mfk_sgd_warp_epoch(&a, st) : plan_.by_row == 2 ? mfk_sgd_cell_epoch(&a, st)
                   : plan_.by_row ? mfk_sgd_run_epoch(&a, st) : mfk_sgd_band_epoch(&a, st));
            step_base_ += (unsigned)plan_.nPass * (unsigned)plan_.nTB;
            launches_++;
            if (world_ > 1) {
                // hand the half-stripe (rows + accumulators) to rank-1, take the one rank+1 just finished
                const int slot = (int)(substeps_done_ % nsub);
                CK(cudaEventRecord((cudaEvent_t)kernel_done_[slot], st));
                CK(cudaStreamWaitEvent(cs, (cudaEvent_t)kernel_done_[slot], 0));
                const size_t rows = (size_t)plan_.stripeRows;
                float *sendS = S0 + (size_t)rs.send_stripe * rows * k_al_, *sendG = SG0 + (size_t)rs.send_stripe * rows * 2;
                float *recvS = S0 + (size_t)rs.recv_stripe * rows * k_al_, *recvG = SG0 + (size_t)rs.recv_stripe * rows * 2;
                NCK(nc->GroupStart());
                NCK(nc->Send(sendS, rows * k_al_, ncclFloat32, rs.send_to, (ncclComm_t)comm_, cs));
                NCK(nc->Send(sendG, rows * 2, ncclFloat32, rs.send_to, (ncclComm_t)comm_, cs));
                NCK(nc->Recv(recvS, rows * k_al_, ncclFloat32, rs.recv_from, (ncclComm_t)comm_, cs));
                NCK(nc->Recv(recvG, rows * 2, ncclFloat32, rs.recv_from, (ncclComm_t)comm_, cs));
                NCK(nc->GroupEnd());
                CK(cudaEventRecord((cudaEvent_t)comm_done_[slot], cs));
                substeps_done_++;
            }
        }
    }
    if (world_ > 1) {
        // all transfers done, then the per-epoch loss sums of all ranks (fpsg_core's table, mf/mf.cpp:2859-2867)
        CK(cudaEventRecord((cudaEvent_t)kernel_done_[0], st));
        CK(cudaStreamWaitEvent(cs, (cudaEvent_t)kernel_done_[0], 0));
        NCK(nc->AllReduce(d_acc_, d_acc_, (size_t)epochs, ncclFloat64, ncclSum, (ncclComm_t)comm_, cs));
        NCK(nc->AllReduce(d_acc_ + kAccErr, d_acc_ + kAccErr, (size_t)epochs, ncclFloat64, ncclSum, (ncclComm_t)comm_, cs));
        CK(cudaEventRecord((cudaEvent_t)comm_done_[0], cs));
        CK(cudaStreamWaitEvent(st, (cudaEvent_t)comm_done_[0], 0));
        gathered_ = false;
    }
    CK(cudaMemcpyAsync(h_acc_, d_acc_, sizeof(double) * (size_t)epochs, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_acc_ + kAccErr, d_acc_ + kAccErr, sizeof(double) * (size_t)epochs, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_acc_ + 1024, d_err_, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (*(int *)(h_acc_ + 1024) != 0) {
        set_error("band schedule wait timed out (code " + std::to_string(*(int *)(h_acc_ + 1024)) + ")");
        return 1;
    }
    const bool hinge = fun_ == MFK_FUN_L2_MFC || fun_ == MFK_FUN_L1_MFC;
    for (int e = 0; e < epochs; e++) {
        loss_out[e] = h_acc_[e];
        err_out[e] = hinge ? h_acc_[kAccErr + e] : h_acc_[e];  // XMMerror = XMMloss for the other losses
    }
    if (d_stats) {
        unsigned long long h[16];
        CK(cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost));
        std::fprintf(stderr,
                     "mfb200 stats (%d epochs): warp iterations %llu, with update %llu, group updates %llu (%.2f per "
                     "updating iteration); idle group-iterations: finished %llu, band not released %llu, no ticket up "
                     "%llu; failed flag polls %llu\n",
                     epochs, h[0], h[1], h[2], h[1] ? (double)h[2] / (double)h[1] : 0.0, h[3], h[4], h[5], h[6]);
        if (plan_.by_row == 2)
            std::fprintf(stderr,
                         "mfb200 stats (cell kernel): runs from the slot %llu, with an exposed load %llu; clock cycles per "
                         "warp iteration with an update %.0f, without %.0f; flag acquisitions %llu, flag raises %llu\n",
                         h[6], h[7], h[1] ? (double)h[8] / (double)h[1] : 0.0,
                         h[0] > h[1] ? (double)h[9] / (double)(h[0] - h[1]) : 0.0, h[10], h[11]);
        dev_free(d_stats);
    }
    return 0;
}

// ---- cross-validation (mf/mf.cpp:3208-3286): hidden grid blocks ------------------------------------------------
void Session::set_hidden_blocks(const int *blocks, int count) {
    hidden_.assign(blocks, blocks + (blocks ? count : 0));
}
bool Session::is_hidden(int blk) const {
    return std::find(hidden_.begin(), hidden_.end(), blk) != hidden_.end();
}
mfk_hidden Session::hidden_arg() const {
    const int bins = std::max(1, prm_.nr_bins);
    mfk_hidden h;
    h.mask = hidden_.empty() ? nullptr : d_hidden_;
    h.bins = bins;
    h.seg_p = std::max(1, (int)std::ceil((double)m_ / bins));  // grid_problem, mf/mf.cpp:799-800
    h.seg_q = std::max(1, (int)std::ceil((double)n_ / bins));
    return h;
}
int Session::upload_hidden_mask() {
    if (hidden_.empty() || d_hidden_) return 0;
    const int bins = std::max(1, prm_.nr_bins), nblk = bins * bins;
    std::vector<unsigned char> mask((size_t)nblk, 0);
    for (int b : hidden_) {
        if (b < 0 || b >= nblk) {
            set_error("hidden block id out of range");
            return 1;
        }
        mask[(size_t)b] = 1;
    }
    if (dev_alloc(&d_hidden_, (size_t)nblk)) return 1;
    CK(cudaMemcpyAsync(d_hidden_, mask.data(), (size_t)nblk, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    CK(cudaStreamSynchronize((cudaStream_t)stream_));
    return 0;
}

// The error measure of the loss over the ratings of the hidden blocks, on the training-space model, scaled back by
// loss family (fpsg_core, mf/mf.cpp:2918-2938 with calc_error 635-674).
int Session::cv_error(double *out) {
    t_pool_stream = (cudaStream_t)stream_;
    cudaStream_t st = (cudaStream_t)stream_;
    *out = 0;
    if (hidden_.empty() || !loaded_) {
        set_error("cv_error: no hidden blocks");
        return 1;
    }
    CK(cudaSetDevice(device_));
    if (gather_model()) return 1;
    if (upload_hidden_mask()) return 1;
    CK(cudaMemsetAsync(d_acc_ + 1034, 0, sizeof(double) * 2, st));
    const float b = avg_ / scale_;
    if (mode_ == MFB200_MODE_EXACT)  // d_R_: training-space ids, ratings already scaled
        CK(mfk_err_general(fun_, d_R_, nnz_, nullptr, nullptr, dP_, dQ_, m_, n_, k_al_, b, 1.0f, d_acc_ + 1034, 1, hidden_arg(), st));
    else  // the raw ratings kept by load_band: ids through the permutations, r * 1/scale
        CK(mfk_err_general(fun_, d_cv_raw_, nnz_, d_pmap_, d_qmap_, dP_, dQ_, m_, n_, k_al_, b, 1.0f / scale_, d_acc_ + 1034, 1,
                           hidden_arg(), st));
    CK(cudaMemcpyAsync(h_acc_ + 1034, d_acc_ + 1034, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double e = h_acc_[1035] > 0 ? h_acc_[1034] / h_acc_[1035] : 0.0;
    if (fun_ == MFK_FUN_L2_MFR)
        e = std::sqrt(e * scale_ * scale_);
    else if (fun_ == MFK_FUN_L1_MFR || fun_ == MFK_FUN_KL_MFR)
        e *= scale_;
    *out = e;
    return 0;
}

// the order in which do_cross_validation deals the grid blocks to the folds (mf/mf.cpp:3210-3214): srand(0) and
// std::random_shuffle over the block ids -- the recurrence of gen_random_map
std::vector<int> cv_block_order(int nr_blocks) { return gen_map(nr_blocks); }

// reg term of the objective column (mf/mf.cpp:2854-2878): (reg1 + reg2) times scale^2 / scale / 1 by loss family.
int Session::objective_terms(double *reg_out) {
    cudaStream_t st = (cudaStream_t)stream_;
    CK(cudaMemsetAsync(d_acc_ + 1030, 0, sizeof(double) * 4, st));
    CK(mfk_reg2(dP_, d_omega_p_, m_, k_al_, d_acc_ + 1030, st));
    CK(mfk_reg2(dQ_, d_omega_q_, n_, k_al_, d_acc_ + 1031, st));
    if (lambda_p1_ != 0) CK(mfk_reg1(dP_, d_omega_p_, m_, k_al_, d_acc_ + 1032, st));
    if (lambda_q1_ != 0) CK(mfk_reg1(dQ_, d_omega_q_, n_, k_al_, d_acc_ + 1033, st));
    CK(cudaMemcpyAsync(h_acc_ + 1030, d_acc_ + 1030, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double reg2 = (double)lambda_p_ * h_acc_[1030] + (double)lambda_q_ * h_acc_[1031];
    const double reg1 = (double)lambda_p1_ * h_acc_[1032] + (double)lambda_q1_ * h_acc_[1033];
    if (fun_ == MFK_FUN_L2_MFR)
        *reg_out = (reg1 + reg2) * scale_ * scale_;
    else if (fun_ == MFK_FUN_L1_MFR || fun_ == MFK_FUN_KL_MFR)
        *reg_out = (reg1 + reg2) * scale_;
    else
        *reg_out = reg1 + reg2;
    return 0;
}

// the va_<metric> column of the table: calc_error / nnz on the training-space model, scaled back by loss family
// (mf/mf.cpp:2884-2904)
int Session::validation_error(double *va_rmse_out) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (gather_model()) return 1;
    CK(cudaMemsetAsync(d_acc_ + 1029, 0, sizeof(double), st));
    if (fun_ == MFK_FUN_L2_MFR)
        CK(mfk_va_err(d_va_, va_nnz_, d_pmap_, d_qmap_, dP_, dQ_, m_, n_, k_al_, avg_ / scale_, 1.0f / scale_, d_acc_ + 1029, st));
    else
        CK(mfk_err_general(fun_, d_va_, va_nnz_, d_pmap_, d_qmap_, dP_, dQ_, m_, n_, k_al_, avg_ / scale_, 1.0f / scale_,
                           d_acc_ + 1029, 1, mfk_hidden{nullptr, 1, 1, 1}, st));
    CK(cudaMemcpyAsync(h_acc_ + 1029, d_acc_ + 1029, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double v = h_acc_[1029] / (double)va_nnz_;
    if (fun_ == MFK_FUN_L2_MFR)
        v = std::sqrt(v * scale_ * scale_);
    else if (fun_ == MFK_FUN_L1_MFR || fun_ == MFK_FUN_KL_MFR)
        v *= scale_;
    *va_rmse_out = v;
    return 0;
}

// the iteration table of fpsg_core, mf/mf.cpp:2818-2832 and 2880-2907 (same widths, precision and
// the same stream state afterwards: cout is left in `scientific`); legends of get_error_legend (745-773).
static const char *error_legend(int fun) {
    switch (fun) {
        case MFK_FUN_L1_MFR: return "mae";
        case MFK_FUN_KL_MFR: return "gkl";
        case MFK_FUN_LR_MFC: return "logloss";
        case MFK_FUN_L2_MFC:
        case MFK_FUN_L1_MFC: return "accuracy";
        case MFK_FUN_ROW_BPR:
        case MFK_FUN_COL_BPR: return "bprloss";
        default: return "rmse";
    }
}
void Session::print_header() {
    if (silent_) return;
    std::cout.width(4);
    std::cout << "iter";
    std::cout.width(13);
    std::cout << std::string("tr_") + error_legend(fun_);
    if (va_nnz_ > 0) {
        std::cout.width(13);
        std::cout << std::string("va_") + error_legend(fun_);
    }
    std::cout.width(13);
    std::cout << "obj";
    std::cout << "\n";
}
void Session::print_row(int iter, double tr_rmse, double va_rmse, double obj) {
    if (silent_) return;
    std::cout.width(4);
    std::cout << iter;
    std::cout.width(13);
    std::cout << std::fixed << std::setprecision(4) << tr_rmse;
    if (va_nnz_ > 0) {
        std::cout.width(13);
        std::cout << std::fixed << std::setprecision(4) << va_rmse;
    }
    std::cout.width(13);
    std::cout << std::fixed << std::setprecision(4) << std::scientific << obj;
    std::cout << "\n" << std::flush;
}

int Session::run_epochs(int epochs, float *ms_out, double *tr_rmse_out, bool print_table) {
    t_pool_stream = (cudaStream_t)stream_;
    if (!loaded_) {
        set_error("session not loaded");
        return 1;
    }
    CK(cudaSetDevice(device_));
    cudaStream_t st = (cudaStream_t)stream_;
    if (nnz_ == 0) {  // mf/mf.cpp:2792-2796
        std::cout << "warning: train on an empty training set" << std::endl;
        if (ms_out) *ms_out = 0.f;
        return 0;
    }
    if (print_table && !header_printed_) {
        print_header();
        header_printed_ = true;
    }
    std::vector<double> loss((size_t)std::max(epochs, 1), 0.0), err((size_t)std::max(epochs, 1), 0.0);
    float ms_total = 0.f;
    // With the table on, epochs are issued one by one (the reference prints after every epoch);
    // quiet runs in ring mode issue all launches back to back.
    const int chunk = (mode_ == MFB200_MODE_RING && !print_table) ? std::min(epochs, 1024) : 1;
    for (int done = 0; done < epochs;) {
        const int now = std::min(chunk, epochs - done);
        CK(cudaEventRecord((cudaEvent_t)ev0_, st));
        if (mode_ == MFB200_MODE_RING) {
            if (epochs_band(now, loss.data() + done, err.data() + done)) return 1;
        } else {
            if (epoch_exact(loss.data() + done, err.data() + done)) return 1;
        }
        CK(cudaEventRecord((cudaEvent_t)ev1_, st));
        CK(cudaEventSynchronize((cudaEvent_t)ev1_));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, (cudaEvent_t)ev0_, (cudaEvent_t)ev1_));
        ms_total += ms;
        for (int e = 0; e < now; e++) {
            // the tr_<metric> and obj columns, mf/mf.cpp:2859-2878
            double tr = err[(size_t)done + e] / (double)nnz_, tr_loss = loss[(size_t)done + e];
            if (fun_ == MFK_FUN_L2_MFR) {
                tr_loss *= scale_ * scale_;
                tr = std::sqrt(tr * scale_ * scale_);
            } else if (fun_ == MFK_FUN_L1_MFR || fun_ == MFK_FUN_KL_MFR) {
                tr_loss *= scale_;
                tr *= scale_;
            }
            last_tr_rmse_ = tr;
            if (tr_rmse_out) tr_rmse_out[done + e] = tr;
            if (print_table) {
                double reg = 0, va = 0;
                if (objective_terms(&reg)) return 1;
                if (va_nnz_ > 0 && validation_error(&va)) return 1;
                last_va_rmse_ = va;
                print_row(epochs_done_ + e, tr, va, reg + tr_loss);
            }
        }
        epochs_done_ += now;
        done += now;
    }
    epochs_ms_ += ms_total;
    if (ms_out) *ms_out = ms_total;
    return 0;
}

// world > 1: after an epoch every rank holds its own T band and its two home half-stripes of S (the
// rotation has come full circle), so two in-place all-gathers rebuild the full model on every rank.
int Session::gather_model() {
    if (world_ <= 1 || gathered_) return 0;
    const NcclApi *nc = nccl_api();
    cudaStream_t st = (cudaStream_t)stream_, cs = (cudaStream_t)comm_stream_;
    const bool sw = plan_.swap_sides != 0;
    float *S = sw ? dP_ : dQ_, *T = sw ? dQ_ : dP_, *SG = sw ? dPG_ : dQG_, *TG = sw ? dQG_ : dPG_;
    const size_t sRows = (size_t)(plan_.nStripes / world_) * plan_.stripeRows, tRows = (size_t)plan_.tSeg;
    CK(cudaEventRecord((cudaEvent_t)kernel_done_[0], st));
    CK(cudaStreamWaitEvent(cs, (cudaEvent_t)kernel_done_[0], 0));
    NCK(nc->GroupStart());
    NCK(nc->AllGather(S + rank_ * sRows * k_al_, S, sRows * k_al_, ncclFloat32, (ncclComm_t)comm_, cs));
    NCK(nc->AllGather(SG + rank_ * sRows * 2, SG, sRows * 2, ncclFloat32, (ncclComm_t)comm_, cs));
    NCK(nc->AllGather(T + rank_ * tRows * k_al_, T, tRows * k_al_, ncclFloat32, (ncclComm_t)comm_, cs));
    NCK(nc->AllGather(TG + rank_ * tRows * 2, TG, tRows * 2, ncclFloat32, (ncclComm_t)comm_, cs));
    NCK(nc->GroupEnd());
    CK(cudaEventRecord((cudaEvent_t)comm_done_[0], cs));
    CK(cudaStreamWaitEvent(st, (cudaEvent_t)comm_done_[0], 0));
    gathered_ = true;
    return 0;
}

int Session::finalize_to_device() {
    cudaStream_t st = (cudaStream_t)stream_;
    if (gather_model()) return 1;
    if (!d_outP_ && dev_alloc(&d_outP_, (size_t)m_ * k_)) return 1;
    if (!d_outQ_ && dev_alloc(&d_outQ_, (size_t)n_ * k_)) return 1;
    const float factor = std::sqrt(scale_);  // scale_model, mf/mf.cpp:551-552
    CK(mfk_finalize_rows(dP_, d_pmap_, m_, k_, k_al_, factor, d_outP_, st));
    CK(mfk_finalize_rows(dQ_, d_qmap_, n_, k_, k_al_, factor, d_outQ_, st));
    return 0;
}

int Session::finish(float *P_out, float *Q_out, float *b_out) {
    t_pool_stream = (cudaStream_t)stream_;
    if (!loaded_) {
        set_error("session not loaded");
        return 1;
    }
    const double t0 = now_ms();
    CK(cudaSetDevice(device_));
    cudaStream_t st = (cudaStream_t)stream_;
    if (finalize_to_device()) return 1;
    if (P_out && staged_d2h(P_out, d_outP_, sizeof(float) * (size_t)m_ * k_, st)) return 1;
    if (Q_out && staged_d2h(Q_out, d_outQ_, sizeof(float) * (size_t)n_ * k_, st)) return 1;
    CK(cudaStreamSynchronize(st));
    if (b_out) {
        float b = avg_ / scale_;  // init_model's b, mf/mf.cpp:3015
        if (scale_ != 1.0f) b *= scale_;  // scale_model, mf/mf.cpp:531-536
        *b_out = b;
    }
    finish_ms_ = now_ms() - t0;
    return 0;
}

int Session::heldout_rmse(const mfb200_node *R, long long nnz, double *out) {
    t_pool_stream = (cudaStream_t)stream_;
    if (!loaded_) {
        set_error("session not loaded");
        return 1;
    }
    CK(cudaSetDevice(device_));
    cudaStream_t st = (cudaStream_t)stream_;
    if (nnz == 0) {
        *out = 0;
        return 0;
    }
    if (finalize_to_device()) return 1;
    mfk_node *d_t = nullptr;
    if (dev_alloc(&d_t, (size_t)nnz)) return 1;
    float b = avg_ / scale_;
    if (scale_ != 1.0f) b *= scale_;
    CK(cudaMemcpyAsync(d_t, R, sizeof(mfk_node) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_acc_ + 1028, 0, sizeof(double), st));
    CK(mfk_sq_err(d_t, nnz, d_outP_, d_outQ_, m_, n_, k_, b, d_acc_ + 1028, st));
    CK(cudaMemcpyAsync(h_acc_ + 1028, d_acc_ + 1028, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dev_free(d_t);
    *out = std::sqrt(h_acc_[1028] / (double)nnz);
    return 0;
}

void Session::fill_report(mfb200_report *r) const {
    std::memset(r, 0, sizeof(*r));
    r->mode_used = mode_;
    r->k_aligned = k_al_;
    if (mode_ == MFB200_MODE_RING) {
        r->grid_ctas = plan_.nC;
        r->cta_warps = plan_.nWarps;
        r->bands = plan_.nTB;
        r->subbands = plan_.nG;
    }
    r->launches = launches_;
    r->prep_ms = prep_ms_;
    r->epochs_ms = epochs_ms_;
    r->finish_ms = finish_ms_;
    r->last_tr_rmse = last_tr_rmse_;
    r->create_ms = create_ms_;
    r->gpus = world_;
    r->kernel = mode_ == MFB200_MODE_RING ? (plan_.by_row == 4 ? 6 : plan_.by_row == 3 ? 4 : plan_.by_row == 2 ? 3 : plan_.by_row ? (plan_.tlock ? 5 : 2) : 1) : 0;
}

}  // namespace mfb200
