// oracle/mf_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the
// product (question-recommendation-system_b200/).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// What it is: a single-threaded CPU restatement of the reference's matrix-
// factorisation hot path as shipped (SSE code path, fun = P_L2_MFR, lambda_1 = 0,
// no NMF), written from SURVEY.md Appendix A.  Every function cites the reference
// lines it restates (paths relative to /root/reference).
//
// Parity status: PINNED.  tests/test_oracle_pinned.py compares this restatement
// bit-for-bit with the reference itself, compiled from /root/reference/mf/mf.cpp
// into oracle/_ref/libmf_ref.so by oracle/Makefile and driven at nr_threads=1
// (the only reproducible mode of the reference), and with the
// golden vectors under tests/golden/ that oracle/make_golden.py produced from
// that compiled reference.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, no -mfma: the reference is
// built with -mavx only, so every multiply and add is rounded separately).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <utility>
#include <vector>
#include <xmmintrin.h>

#include "rsqrt12_table.h"

extern "C" {

struct orc_node {  // mf/mf.h:36-41 (mf_node), same layout
    int u;
    int v;
    float r;
};

struct orc_param {  // the subset of mf_parameter (mf/mf.h:51-66) the path can vary
    int k;
    int nr_bins;
    int nr_iters;
    float lambda_p2;
    float lambda_q2;
    float eta;
    int rsqrt_mode;  // 0 = 2048-entry table (portable, == Intel hardware), 1 = this CPU's RSQRTSS
};

// orc_param + the remaining knobs of mf_parameter that change the arithmetic of the path (mf/mf.h:51-66):
// loss function, L1 regularisation, non-negativity.  fun uses the reference's codes (mf/mf.h:25-33):
// 0 P_L2_MFR, 1 P_L1_MFR, 2 P_KL_MFR, 5 P_LR_MFC, 6 P_L2_MFC, 7 P_L1_MFC.
struct orc_param_ex {
    orc_param base;
    int fun;          // also 10 P_ROW_BPR_MFOC, 11 P_COL_BPR_MFOC (one-class BPR, mf/mf.cpp:2131-2707)
    float lambda_p1;
    float lambda_q1;
    int do_nmf;
    unsigned rand_seed;  // BPR only: the process-wide rand() is in the state srand(rand_seed) leaves when training is
                         // entered (a fresh process: 1); the scheduler seeds its per-block generators from it (109)
};

}  // extern "C"

namespace {

// ---- library-defined random sequences the reference depends on -------------------

// std::default_random_engine == minstd_rand0 (x <- 16807 x mod 2^31-1, seed 1) feeding
// uniform_real_distribution<float>(0,1): one draw, value float(x-1)/2^31, clamped below 1.
// Used at mf/mf.cpp:82-83,108,203 (scheduler priorities) and 972-973,995 (initial factors).
struct Minstd {
    uint32_t x = 1;
    float u01() {
        x = (uint32_t)(((uint64_t)x * 16807u) % 2147483647u);
        float f = (float)(x - 1u) / 2147483648.0f;
        if (f >= 1.0f) f = std::nextafter(1.0f, 0.0f);
        return f;
    }
};

// glibc rand() after srand(seed) (TYPE_3 additive feedback generator, degree 31, sep 3).
// Used through std::random_shuffle at mf/mf.cpp:1011-1015.
struct GlibcRand {
    int32_t ring[32];
    uint32_t pos;  // index of the next value to produce
    explicit GlibcRand(uint32_t seed) {
        std::vector<int32_t> r(344);
        if (seed == 0) seed = 1;
        r[0] = (int32_t)seed;
        for (int i = 1; i < 31; i++) {
            int64_t w = (16807LL * r[i - 1]) % 2147483647LL;
            if (w < 0) w += 2147483647LL;
            r[i] = (int32_t)w;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 34; i < 344; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
        for (int i = 344 - 32; i < 344; i++) ring[i & 31] = r[i];
        pos = 344;
    }
    int next() {
        uint32_t o = (uint32_t)ring[(pos - 31) & 31] + (uint32_t)ring[(pos - 3) & 31];
        ring[pos & 31] = (int32_t)o;
        pos++;
        return (int)(o >> 1);
    }
};

// gen_random_map, mf/mf.cpp:1009-1017: srand(0); identity; std::random_shuffle, which in
// libstdc++ is: for i = 1..size-1: j = rand() % (i+1); swap(a[i], a[j]).
std::vector<int> random_map(int size) {
    GlibcRand g(0);
    std::vector<int> a(size);
    for (int i = 0; i < size; i++) a[i] = i;
    for (int i = 1; i < size; i++) {
        int j = g.next() % (i + 1);
        if (i != j) std::swap(a[i], a[j]);
    }
    return a;
}

// ---- the approximate reciprocal square root (mf/mf.cpp:1469-1470) -----------------

float rsqrt12_table(float x) {
    uint32_t b;
    memcpy(&b, &x, 4);
    uint32_t e = (b >> 23) & 0xffu, man = b & 0x7fffffu;
    uint32_t out;
    if (e == 0xffu)
        out = man ? (b | 0x400000u) : ((b >> 31) ? 0xffc00000u : 0u);  // NaN -> qNaN ; +inf -> 0 ; -inf -> NaN
    else if (e == 0u)
        out = (b & 0x80000000u) | 0x7f800000u;  // +-0 and (DAZ-style) denormals -> +-inf
    else if (b >> 31)
        out = 0xffc00000u;  // negative -> default NaN
    else {
        uint32_t p = (e & 1u) ? 0u : 1u;
        int32_t sh = ((int32_t)e - (127 + (int32_t)p)) / 2;
        out = ORC_RSQRT12_TABLE[p * 1024u + (man >> 13)] - ((uint32_t)sh << 23);
    }
    float y;
    memcpy(&y, &out, 4);
    return y;
}

float rsqrt12_hw(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }

// ---- 4-lane arithmetic in the order of the SSE code path --------------------------

// calc_z, mf/mf.cpp:1264-1273 (also Utility::inner_product, 557-566): lane j accumulates
// dims d = j (mod 4) in ascending order, product rounded before the add; then two HADDs.
inline float dot_sse_order(const float *p, const float *q, int k_al) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int d = 0; d < k_al; d += 4) {
        l0 = l0 + p[d] * q[d];
        l1 = l1 + p[d + 1] * q[d + 1];
        l2 = l2 + p[d + 2] * q[d + 2];
        l3 = l3 + p[d + 3] * q[d + 3];
    }
    return (l0 + l1) + (l2 + l3);
}

struct Solver {
    float lambda_p, lambda_q, eta;
    float (*rsq)(float);

    // MFSolver::sg_update, mf/mf.cpp:1462-1548 with lambda_1 == 0 and do_nmf == false.
    // `e` is the error r - z (the reference keeps it in the register named XMMz).
    // rk is 1/8 for BOTH halves in the shipped SSE path (mf/mf.cpp:1228-1234).
    inline void half_update(float *p, float *q, float *pG, float *qG, float e, int d0, int d1) const {
        const float eta_p = eta * rsq(*pG);
        const float eta_q = eta * rsq(*qG);
        float sp[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
        for (int d = d0; d < d1; d += 4)
            for (int j = 0; j < 4; j++) {
                const float pv = p[d + j], qv = q[d + j];
                const float gp = lambda_p * pv - e * qv;
                const float gq = lambda_q * qv - e * pv;
                sp[j] = sp[j] + gp * gp;
                sq[j] = sq[j] + gq * gq;
                p[d + j] = pv - eta_p * gp;
                q[d + j] = qv - eta_q * gq;
            }
        *pG = *pG + ((sp[0] + sp[1]) + (sp[2] + sp[3])) * 0.125f;
        *qG = *qG + ((sq[0] + sq[1]) + (sq[2] + sq[3])) * 0.125f;
    }
};

// The six prepare_for_sg_update variants of the SSE code path (mf/mf.cpp:1719-1728, 1768-1781, 1829-1840,
// 1884-1903, 1965-1988, 2053-2080): from z = <p,q> and the rating to the scalar the update multiplies the
// other row with (the reference keeps it in XMMz), plus the two per-block accumulators.  `loss` and `error`
// are the doubles the float results are widened into.
inline float prepare_scalar(int fun, float z, float r, double &loss, double &error) {
    switch (fun) {
        case 0: {  // P_L2_MFR
            z = r - z;
            loss += (double)(z * z);
            error = loss;
            return z;
        }
        case 1: {  // P_L1_MFR
            z = r - z;
            loss += (double)std::fabs(z);
            error = loss;
            return (z > 0.0f ? 1.0f : 0.0f) + (z < 0.0f ? -1.0f : 0.0f);
        }
        case 2: {  // P_KL_MFR: log is the float overload (float argument, `using namespace std`)
            z = r / z;
            loss += (double)(float)(r * (std::log(z) - 1 + 1 / z));
            error = loss;
            return z - 1.0f;
        }
        case 5: {  // P_LR_MFC: exp and log are the float overloads; the loss is widened after the log
            if (r > 0) {
                z = std::exp(-z);
                loss += (double)std::log(1 + z);
                error = loss;
                return z / (1 + z);
            }
            z = std::exp(z);
            loss += (double)std::log(1 + z);
            error = loss;
            return -z / (1 + z);
        }
        case 6: {  // P_L2_MFC: error counts the correctly classified ratings
            if (r > 0) {
                error += (double)(z > 0.0f ? 1.0f : 0.0f);
                const float t = 1.0f - z;
                z = 0.0f > t ? 0.0f : t;  // _mm_max_ps(0, t): second operand unless the first is greater
            } else {
                error += (double)(z < 0.0f ? 1.0f : 0.0f);
                const float t = -1.0f - z;
                z = 0.0f < t ? 0.0f : t;  // _mm_min_ps(0, t)
            }
            loss += (double)(z * z);
            return z;
        }
        case 7: {  // P_L1_MFC
            if (r > 0) {
                error += (double)(z >= 0.0f ? 1.0f : 0.0f);
                z = 1.0f - z;
                loss += (double)(0.0f > z ? 0.0f : z);
                return z >= 0.0f ? 1.0f : 0.0f;
            }
            error += (double)(z < 0.0f ? 1.0f : 0.0f);
            z = 1.0f + z;
            loss += (double)(0.0f > z ? 0.0f : z);
            return z >= 0.0f ? -1.0f : 0.0f;
        }
    }
    return 0.0f;
}

// MFSolver::sg_update (mf/mf.cpp:1462-1548) in full: the L2 step, then -- each as a separate pass over the
// half, as the reference does -- the L1 soft threshold (1499-1527) and the projection on the non-negative
// orthant (1529-1541).  The AdaGrad sums use the L2 gradient only.
struct SolverEx {
    float lambda_p1, lambda_q1, lambda_p2, lambda_q2, eta;
    bool do_nmf;
    float (*rsq)(float);

    static inline float soft(float x, float step) {
        // flip = (x <= 0) ? -0.0 : +0.0; x = flip ^ max((x ^ flip) - step, 0)
        uint32_t b;
        memcpy(&b, &x, 4);
        const uint32_t flip = (x <= 0.0f) ? 0x80000000u : 0u;
        b ^= flip;
        float a;
        memcpy(&a, &b, 4);
        a = a - step;
        a = a > 0.0f ? a : 0.0f;  // _mm_max_ps(a, 0): 0 unless a is greater (also for NaN)
        memcpy(&b, &a, 4);
        b ^= flip;
        memcpy(&a, &b, 4);
        return a;
    }

    inline void half_update(float *p, float *q, float *pG, float *qG, float zz, int d0, int d1) const {
        const float eta_p = eta * rsq(*pG);
        const float eta_q = eta * rsq(*qG);
        float sp[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
        for (int d = d0; d < d1; d += 4)
            for (int j = 0; j < 4; j++) {
                const float pv = p[d + j], qv = q[d + j];
                const float gp = lambda_p2 * pv - zz * qv;
                const float gq = lambda_q2 * qv - zz * pv;
                sp[j] = sp[j] + gp * gp;
                sq[j] = sq[j] + gq * gq;
                p[d + j] = pv - eta_p * gp;
                q[d + j] = qv - eta_q * gq;
            }
        if (lambda_p1 > 0) {
            const float step = eta_p * lambda_p1;
            for (int d = d0; d < d1; d++) p[d] = soft(p[d], step);
        }
        if (lambda_q1 > 0) {
            const float step = eta_q * lambda_q1;
            for (int d = d0; d < d1; d++) q[d] = soft(q[d], step);
        }
        if (do_nmf)
            for (int d = d0; d < d1; d++) {
                p[d] = p[d] > 0.0f ? p[d] : 0.0f;  // _mm_max_ps(x, 0)
                q[d] = q[d] > 0.0f ? q[d] : 0.0f;
            }
        *pG = *pG + ((sp[0] + sp[1]) + (sp[2] + sp[3])) * 0.125f;
        *qG = *qG + ((sq[0] + sq[1]) + (sq[2] + sq[3])) * 0.125f;
    }
};

// BPRSolver (mf/mf.cpp:2131-2335, SSE path): z = <p, q - w> in the 4-lane order (2182-2191), the scalar
// exp(-z) / (1 + exp(-z)) with the float exp (2325-2335), then per half the three-row step of 2211-2323: gradients
// from the OLD rows, AdaGrad sums with rk = 1/8 for both halves (run() passes rk_slow twice, 1228-1234), the L1
// threshold (lambda_p1 on p, lambda_q1 on q AND w) and the projection as separate passes.
struct SolverBpr {
    float lambda_p1, lambda_q1, lambda_p2, lambda_q2, eta;
    bool do_nmf;
    float (*rsq)(float);

    static inline float dot3(const float *p, const float *q, const float *w, int k_al) {
        float lane[4] = {0.f, 0.f, 0.f, 0.f};
        for (int d = 0; d < k_al; d += 4)
            for (int j = 0; j < 4; j++) lane[j] = lane[j] + p[d + j] * (q[d + j] - w[d + j]);
        return (lane[0] + lane[1]) + (lane[2] + lane[3]);
    }
    inline void half_update(float *p, float *q, float *w, float *pG, float *qG, float *wG, float zz, int d0, int d1) const {
        // the negative row may BE the positive row (w == q): the reference loads the three accumulators first and stores
        // p, q, w in this order, so the w results win -- keep its loads and stores in its order
        const float pG0 = *pG, qG0 = *qG, wG0 = *wG;
        const float eta_p = eta * rsq(pG0), eta_q = eta * rsq(qG0), eta_w = eta * rsq(wG0);
        float sp[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f}, sw[4] = {0.f, 0.f, 0.f, 0.f};
        for (int d = d0; d < d1; d += 4)
            for (int j = 0; j < 4; j++) {
                const float pv = p[d + j], qv = q[d + j], wv = w[d + j];
                const float pg = lambda_p2 * pv + zz * (wv - qv);
                const float qg = lambda_q2 * qv - zz * pv;
                const float wg = lambda_q2 * wv + zz * pv;
                sp[j] = sp[j] + pg * pg;
                sq[j] = sq[j] + qg * qg;
                sw[j] = sw[j] + wg * wg;
                p[d + j] = pv - eta_p * pg;
                q[d + j] = qv - eta_q * qg;
                w[d + j] = wv - eta_w * wg;
            }
        if (lambda_p1 > 0) {
            const float step = eta_p * lambda_p1;
            for (int d = d0; d < d1; d++) p[d] = SolverEx::soft(p[d], step);
        }
        if (lambda_q1 > 0) {
            const float sq1 = eta_q * lambda_q1, sw1 = eta_w * lambda_q1;
            for (int d = d0; d < d1; d += 4) {  // 2291-2309: q and w of four dimensions are loaded, then q is stored, then w
                float qv[4], wv[4];
                for (int j = 0; j < 4; j++) {
                    qv[j] = q[d + j];
                    wv[j] = w[d + j];
                }
                for (int j = 0; j < 4; j++) q[d + j] = SolverEx::soft(qv[j], sq1);
                for (int j = 0; j < 4; j++) w[d + j] = SolverEx::soft(wv[j], sw1);
            }
        }
        if (do_nmf)
            for (int d = d0; d < d1; d++) {
                p[d] = p[d] > 0.0f ? p[d] : 0.0f;
                q[d] = q[d] > 0.0f ? q[d] : 0.0f;
                w[d] = w[d] > 0.0f ? w[d] : 0.0f;
            }
        *pG = pG0 + ((sp[0] + sp[1]) + (sp[2] + sp[3])) * 0.125f;
        *qG = qG0 + ((sq[0] + sq[1]) + (sq[2] + sq[3])) * 0.125f;
        *wG = wG0 + ((sw[0] + sw[1]) + (sw[2] + sw[3])) * 0.125f;
    }
};

// Scheduler::get_negative (mf/mf.cpp:249-280): one draw of the first block's minstd_rand0; odd -> a row of the first
// block's range, even -> of the second block's
inline int bpr_negative(uint32_t &gen_state, int first_block, int second_block, int m, int n, int bins, bool col) {
    gen_state = (uint32_t)(((uint64_t)gen_state * 16807u) % 2147483647u);
    const int rand_val = (int)gen_state;
    auto gen_random = [&](int block_id) {
        int v_min, v_max;
        if (col) {
            const int seg = (int)std::ceil((double)m / bins);
            v_min = std::min((block_id / bins) * seg, m - 1);
            v_max = std::min(v_min + seg, m - 1);
        } else {
            const int seg = (int)std::ceil((double)n / bins);
            v_min = std::min((block_id % bins) * seg, n - 1);
            v_max = std::min(v_min + seg, n - 1);
        }
        return v_max == v_min ? v_min : rand_val % (v_max - v_min) + v_min;
    };
    return (rand_val % 2) ? gen_random(first_block) : gen_random(second_block);
}

struct ByUV {
    bool operator()(const orc_node &a, const orc_node &b) const {
        return a.u != b.u ? a.u < b.u : a.v < b.v;
    }
};
struct ByVU {
    bool operator()(const orc_node &a, const orc_node &b) const {
        return a.v != b.v ? a.v < b.v : a.u < b.u;
    }
};

struct FtzScope {  // fpsg_core turns flush-to-zero on for the whole loop, mf/mf.cpp:2788-2791,2940-2942
    unsigned old;
    FtzScope() : old(_MM_GET_FLUSH_ZERO_MODE()) { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); }
    ~FtzScope() { _MM_SET_FLUSH_ZERO_MODE(old); }
};

inline float predict_one(const float *P, const float *Q, int m, int n, int k, float b, int u, int v) {
    // mf_predict, mf/mf.cpp:4295-4314.  The sign-threshold branch at 4308-4311 is dead (&& chain).
    if (u < 0 || u >= m || v < 0 || v >= n) return b;
    const float *p = P + (long long)u * k, *q = Q + (long long)v * k;
    float z = 0.0f;
    for (int d = 0; d < k; d++) z = z + p[d] * q[d];  // std::inner_product order, no FMA
    if (std::isnan(z)) z = b;
    return z;
}

}  // namespace

extern "C" {

// Whole training pipeline at nr_threads = 1: fpsg + fpsg_core, mf/mf.cpp:2945-3042,2774-2943
// (SURVEY.md Appendix A.2 steps 0-10).  Outputs: P[m*k], Q[n*k] (stride k, original ids), *b,
// tr_rmse[nr_iters] and obj[nr_iters] (the two numbers of the per-iteration table, 2852-2907).
// Returns 0, or 1 for an empty training set (2792-2796: model stays as initialised).
int orc_train_ex(const orc_node *R_in, long long nnz, int m, int n, const orc_param_ex *prx, float *P_out,
                 float *Q_out, float *b_out, double *tr_rmse, double *obj);

int orc_train(const orc_node *R_in, long long nnz, int m, int n, const orc_param *prm, float *P_out,
              float *Q_out, float *b_out, double *tr_rmse, double *obj) {
    orc_param_ex ex;
    ex.base = *prm;
    ex.fun = 0;
    ex.lambda_p1 = ex.lambda_q1 = 0.0f;
    ex.do_nmf = 0;
    return orc_train_ex(R_in, nnz, m, n, &ex, P_out, Q_out, b_out, tr_rmse, obj);
}

// The same pipeline for every MFSolver loss (fun 0,1,2,5,6,7), with L1 regularisation and NMF.  tr_rmse
// receives the tr_<metric> column of the table (rmse / mae / gkl / logloss / accuracy), obj the obj column.
static int train_core(const orc_node *R_in, long long nnz, int m, int n, const orc_param_ex *prx, float *P_out,
                      float *Q_out, float *b_out, double *tr_rmse, double *obj, const int *hidden, int nhidden,
                      double *cv_error);

int orc_train_ex(const orc_node *R_in, long long nnz, int m, int n, const orc_param_ex *prx, float *P_out,
                 float *Q_out, float *b_out, double *tr_rmse, double *obj) {
    return train_core(R_in, nnz, m, n, prx, P_out, Q_out, b_out, tr_rmse, obj, nullptr, 0, nullptr);
}

// One fold of the cross-validation (CrossValidator::do_cv1 -> fpsg with cv_blocks, mf/mf.cpp:3281-3286): the grid blocks
// listed in `hidden` are never scheduled (Scheduler constructor, 104-111: they get no priority and no draw of the
// engine), an "epoch" is still nr_bins^2 finished jobs (target, 94 and 305), and after training the error measure of
// the loss is taken over the hidden blocks on the training-space model (fpsg_core, 2918-2938).
int orc_train_cv(const orc_node *R_in, long long nnz, int m, int n, const orc_param_ex *prx, const int *hidden,
                 int nhidden, double *cv_error) {
    std::vector<float> P((size_t)m * prx->base.k), Q((size_t)n * prx->base.k);
    float b;
    return train_core(R_in, nnz, m, n, prx, P.data(), Q.data(), &b, nullptr, nullptr, hidden, nhidden, cv_error);
}

// mf_cross_validation (mf/mf.cpp:4117-4129) = CrossValidatorBase::do_cross_validation (3208-3262): srand(0), the block
// ids shuffled with std::random_shuffle (the same recurrence as gen_random_map), fold f hides blocks
// [f*bpf, min((f+1)*bpf, nblk)) of that order with bpf = nblk / nr_folds; returns the mean of the fold errors.
double orc_cross_validation(const orc_node *R_in, long long nnz, int m, int n, const orc_param_ex *prx, int nr_folds,
                            double *fold_errors) {
    const int nblk = prx->base.nr_bins * prx->base.nr_bins, bpf = nblk / nr_folds;
    const std::vector<int> order = random_map(nblk);
    double sum = 0;
    for (int f = 0; f < nr_folds; f++) {
        const int lo = f * bpf, hi = std::min((f + 1) * bpf, nblk);
        double err = 0;
        train_core(R_in, nnz, m, n, prx, nullptr, nullptr, nullptr, nullptr, nullptr, order.data() + lo, hi - lo, &err);
        if (fold_errors) fold_errors[f] = err;
        sum += err;
    }
    return sum / nr_folds;
}

static int train_core(const orc_node *R_in, long long nnz, int m, int n, const orc_param_ex *prx, float *P_out,
                      float *Q_out, float *b_out, double *tr_rmse, double *obj, const int *hidden, int nhidden,
                      double *cv_error) {
    const orc_param *prm = &prx->base;
    const int fun = prx->fun;
    const bool regression = fun == 0 || fun == 1 || fun == 2;
    const int bins = prm->nr_bins, k = prm->k, nblk = bins * bins;

    // step 0: the scheduler's own engine draws bins^2 initial priorities first (89-111).
    Minstd sched_rng;
    typedef std::pair<float, int> Job;
    std::priority_queue<Job, std::vector<Job>, std::greater<Job>> heap;
    std::vector<char> is_hidden(nblk, 0);
    for (int i = 0; i < nhidden; i++) is_hidden[hidden[i]] = 1;
    const bool bpr = fun == 10 || fun == 11, col = fun == 11;
    // (103-110: one loop draws a block's priority and seeds its minstd_rand0 from the process-wide rand())
    GlibcRand proc_rand(prx->rand_seed ? prx->rand_seed : 1u);
    std::vector<uint32_t> block_gen(nblk, 1u);
    for (int i = 0; i < nblk; i++) {
        if (!is_hidden[i]) heap.push(Job(sched_rng.u01(), i));
        if (bpr) {
            uint32_t sd = (uint32_t)proc_rand.next() % 2147483647u;
            block_gen[i] = sd ? sd : 1u;
        }
    }
    if (bpr && (nhidden > 0 || cv_error)) return 2;  // cross-validation of the BPR losses draws from rand(): not restated
    std::vector<int> visits(nblk, 0);

    // step 1-2: collect_info (462-484), scale (2996-2999).
    std::vector<orc_node> R(R_in, R_in + nnz);
    double ex = 0, ex2 = 0;
    for (long long i = 0; i < nnz; i++) {
        ex += (double)R[i].r;
        ex2 += (double)R[i].r * R[i].r;
    }
    ex /= (double)nnz;
    ex2 /= (double)nnz;
    const float avg = (float)ex, std_dev = (float)std::sqrt(ex2 - ex * ex);
    const float scale = regression ? std::max(1e-4f, std_dev) : 1.0f;  // 2996-2999: only the regression losses scale

    // step 3-4: permutations, remap, rating scaling (1009-1017, 775-791, 517-527, 3008-3011).
    std::vector<int> p_map = random_map(m), q_map = random_map(n);
    const float inv_scale = 1.0f / scale;
    for (long long i = 0; i < nnz; i++) {
        R[i].u = p_map[R[i].u];
        R[i].v = q_map[R[i].v];
        if (inv_scale != 1.0f) R[i].r *= inv_scale;
    }

    // step 5: grid_problem (793-858): counts, in-place bucket permutation, per-block sort.
    const int seg_p = (int)std::ceil((double)m / bins), seg_q = (int)std::ceil((double)n / bins);
    std::vector<long long> first(nblk + 1, 0);
    std::vector<int> omega_p(m, 0), omega_q(n, 0);
    {
        std::vector<long long> cnt(nblk, 0);
        for (long long i = 0; i < nnz; i++) {
            cnt[(R[i].u / seg_p) * bins + R[i].v / seg_q]++;
            omega_p[R[i].u]++;
            omega_q[R[i].v]++;
        }
        for (int b = 0; b < nblk; b++) first[b + 1] = first[b] + cnt[b];
        std::vector<long long> fill(first.begin(), first.end() - 1);
        for (int b = 0; b < nblk; b++)
            for (long long at = fill[b]; at != first[b + 1];) {
                const int home = (R[at].u / seg_p) * bins + R[at].v / seg_q;
                if (home == b)
                    at++;
                else
                    std::swap(R[at], R[fill[home]++]);
            }
        for (int b = 0; b < nblk; b++) {
            if (m > n)
                std::sort(R.begin() + first[b], R.begin() + first[b + 1], ByUV());
            else
                std::sort(R.begin() + first[b], R.begin() + first[b + 1], ByVU());
        }
    }

    // step 6: init_model (952-1007): k_al, one engine for P then Q, NaN rows for unseen ids.
    const int k_al = ((k + 7) / 8) * 8;
    std::vector<float> P((size_t)m * k_al, 0.f), Q((size_t)n * k_al, 0.f);
    {
        Minstd init_rng;
        const float s = (float)std::sqrt(1.0 / k);
        for (int pass = 0; pass < 2; pass++) {
            std::vector<float> &M = pass ? Q : P;
            const std::vector<int> &om = pass ? omega_q : omega_p;
            const int rows = pass ? n : m;
            for (int i = 0; i < rows; i++)
                for (int d = 0; d < k; d++)
                    M[(size_t)i * k_al + d] =
                        om[i] > 0 ? (float)(init_rng.u01() * s) : bpr ? 0.0f : std::numeric_limits<float>::quiet_NaN();  // 997
        }
    }
    float b = avg / scale;
    if (nnz == 0) {
        if (b_out) *b_out = b;
        return 1;
    }

    // step 7-8: fpsg_core (2774-2943) and SolverBase::run (1201-1238) for one thread.
    SolverEx sv;
    sv.lambda_p2 = prm->lambda_p2;
    sv.lambda_q2 = prm->lambda_q2;
    sv.lambda_p1 = prx->lambda_p1;
    sv.lambda_q1 = prx->lambda_q1;
    if (fun == 0) {  // 2798-2816: the regularisation coefficients follow the rating scale
        sv.lambda_p2 /= scale;
        sv.lambda_q2 /= scale;
        sv.lambda_p1 /= (float)std::pow(scale, 1.5);
        sv.lambda_q1 /= (float)std::pow(scale, 1.5);
    } else if (fun == 1 || fun == 2) {
        sv.lambda_p1 /= std::sqrt(scale);
        sv.lambda_q1 /= std::sqrt(scale);
    }
    sv.eta = prm->eta;
    sv.do_nmf = prx->do_nmf != 0;
    sv.rsq = prm->rsqrt_mode == 1 ? rsqrt12_hw : rsqrt12_table;
    std::vector<float> PG((size_t)m * 2, 1.f), QG((size_t)n * 2, 1.f);
    std::vector<double> blk_loss(nblk, 0.0), blk_error(nblk, 0.0);
    const bool l1_free = sv.lambda_p1 == 0 && sv.lambda_q1 == 0;  // 2834
    {
        FtzScope ftz;
        for (int it = 0; it < prm->nr_iters; it++) {
            const bool slow_only = l1_free && (it == 0);  // 2834 + 2910-2911
            for (int job = 0; job < nblk; job++) {
                const Job top = heap.top();
                heap.pop();
                const int blk = top.second;
                visits[blk]++;
                double loss = 0.0, error = 0.0;  // arrange_block zeroes both per block, 1254-1262
                if (bpr) {
                    // get_bpr_job (152-191) at one thread: the first block in priority order that shares blk's row band
                    // (column band when column-oriented) and not its other band; it stays out of the queue meanwhile
                    int second = blk;
                    {
                        std::vector<Job> locked;
                        while (!heap.empty()) {
                            const Job cand = heap.top();
                            heap.pop();
                            const int pb = cand.second / bins, qb = cand.second % bins;
                            const bool rejected = col ? (blk % bins != qb || pb == blk / bins) : (blk / bins != pb || qb == blk % bins);
                            if (rejected) {
                                locked.push_back(cand);
                            } else {
                                second = cand.second;
                                break;
                            }
                        }
                        for (const Job &j : locked) heap.push(j);
                    }
                    SolverBpr sb;
                    sb.eta = sv.eta; sb.do_nmf = sv.do_nmf; sb.rsq = sv.rsq;
                    // COL_BPR_MFOC::load_fixed_variables swaps the coefficients with the rows (2645-2686)
                    sb.lambda_p1 = col ? sv.lambda_q1 : sv.lambda_p1;
                    sb.lambda_q1 = col ? sv.lambda_p1 : sv.lambda_q1;
                    sb.lambda_p2 = col ? sv.lambda_q2 : sv.lambda_p2;
                    sb.lambda_q2 = col ? sv.lambda_p2 : sv.lambda_q2;
                    for (long long i = first[blk]; i < first[blk + 1]; i++) {
                        const orc_node &N = R[i];
                        const int neg = bpr_negative(block_gen[blk], blk, second, m, n, bins, col);
                        // row-oriented: p = user, q = item, w = another item; column-oriented: p = item, q = user, w = another user
                        float *p = col ? &Q[(size_t)N.v * k_al] : &P[(size_t)N.u * k_al];
                        float *q = col ? &P[(size_t)N.u * k_al] : &Q[(size_t)N.v * k_al];
                        float *w = col ? &P[(size_t)neg * k_al] : &Q[(size_t)neg * k_al];
                        float *pG = col ? &QG[(size_t)N.v * 2] : &PG[(size_t)N.u * 2];
                        float *qG = col ? &PG[(size_t)N.u * 2] : &QG[(size_t)N.v * 2];
                        float *wG = col ? &PG[(size_t)neg * 2] : &QG[(size_t)neg * 2];
                        float z = SolverBpr::dot3(p, q, w, k_al);
                        z = std::exp(-z);                     // float overload, 2331
                        loss += (double)std::log(1 + z);      // float log widened, 2332
                        error = loss;
                        z = z / (1 + z);
                        sb.half_update(p, q, w, pG, qG, wG, z, 0, 8);
                        if (!slow_only) sb.half_update(p, q, w, pG + 1, qG + 1, wG + 1, z, 8, k_al);
                    }
                    blk_loss[blk] = loss;
                    blk_error[blk] = error;
                    heap.push(Job((float)visits[blk] + sched_rng.u01(), blk));              // put_job, 202-204
                    if (second != blk) heap.push(Job((float)visits[second] + sched_rng.u01(), second));  // put_bpr_job, 222-235
                    continue;
                }
                for (long long i = first[blk]; i < first[blk + 1]; i++) {
                    const orc_node &N = R[i];
                    float *p = &P[(size_t)N.u * k_al], *q = &Q[(size_t)N.v * k_al];
                    const float e = prepare_scalar(fun, dot_sse_order(p, q, k_al), N.r, loss, error);
                    sv.half_update(p, q, &PG[(size_t)N.u * 2], &QG[(size_t)N.v * 2], e, 0, 8);
                    if (!slow_only)
                        sv.half_update(p, q, &PG[(size_t)N.u * 2 + 1], &QG[(size_t)N.v * 2 + 1], e, 8, k_al);
                }
                blk_loss[blk] = loss;
                blk_error[blk] = error;
                heap.push(Job((float)visits[blk] + sched_rng.u01(), blk));  // 202-204
            }
            // the numbers of the iteration table (2854-2878)
            double tr_loss = 0.0, tr_error = 0.0;
            for (int bb = 0; bb < nblk; bb++) tr_loss += blk_loss[bb];
            for (int bb = 0; bb < nblk; bb++) tr_error += blk_error[bb];
            tr_error /= (double)nnz;
            // calc_reg1 (583-606): float sum of |x| over all k_al dims of the seen rows, times omega, in double
            auto reg1_core = [&](const std::vector<float> &M, int rows, const std::vector<int> &om) {
                double reg = 0;
                for (int i = 0; i < rows; i++) {
                    if (om[i] <= 0) continue;
                    float tmp = 0;
                    for (int j = 0; j < k_al; j++) tmp += std::fabs(M[(size_t)i * k_al + j]);
                    reg += om[i] * tmp;
                }
                return reg;
            };
            const double reg1 = sv.lambda_p1 * reg1_core(P, m, omega_p) + sv.lambda_q1 * reg1_core(Q, n, omega_q);
            double reg_p = 0.0, reg_q = 0.0;
            for (int i = 0; i < m; i++)
                if (omega_p[i] > 0) {
                    const float *row = &P[(size_t)i * k_al];
                    reg_p += omega_p[i] * dot_sse_order(row, row, k_al);  // int * float, 608-633
                }
            for (int i = 0; i < n; i++)
                if (omega_q[i] > 0) {
                    const float *row = &Q[(size_t)i * k_al];
                    reg_q += omega_q[i] * dot_sse_order(row, row, k_al);
                }
            const double reg2 = sv.lambda_p2 * reg_p + sv.lambda_q2 * reg_q;
            double reg;
            if (fun == 0) {
                reg = (reg1 + reg2) * scale * scale;
                tr_loss *= scale * scale;
                tr_error = std::sqrt(tr_error * scale * scale);
            } else if (fun == 1 || fun == 2) {
                reg = (reg1 + reg2) * scale;
                tr_loss *= scale;
                tr_error *= scale;
            } else {
                reg = reg1 + reg2;
            }
            if (tr_rmse) tr_rmse[it] = tr_error;
            if (obj) obj[it] = reg + tr_loss;
        }
    }

    // cross-validation error over the hidden blocks, training-space model (2918-2938, calc_error 635-674)
    if (cv_error && nhidden > 0) {
        double err = 0;
        long long cv_count = 0;
        for (int h = 0; h < nhidden; h++) {
            const int blk = hidden[h];
            cv_count += first[blk + 1] - first[blk];
            for (long long i = first[blk]; i < first[blk + 1]; i++) {
                const orc_node &N = R[i];
                const float z = predict_one(P.data(), Q.data(), m, n, k_al, b, N.u, N.v);
                switch (fun) {
                    case 0: err += std::pow((double)(N.r - z), 2); break;
                    case 1: err += std::fabs(N.r - z); break;
                    case 2: err += N.r * std::log(N.r / z) - N.r + z; break;
                    case 5: err += N.r > 0 ? std::log(1.0 + std::exp(-z)) : std::log(1.0 + std::exp(z)); break;
                    default: err += N.r > 0 ? (z > 0 ? 1 : 0) : (z < 0 ? 1 : 0); break;
                }
            }
        }
        err /= (double)cv_count;
        if (fun == 0)
            err = std::sqrt(err * scale * scale);
        else if (fun == 1 || fun == 2)
            err *= scale;
        *cv_error = err;
    }
    if (!P_out) return 0;

    // step 9-10: scale_model over all k_al dims (529-553), shrink (1057-1074), un-permute (1027-1055).
    b *= scale;
    const float fs = std::sqrt(scale);
    if (scale != 1.0f) {
        for (size_t i = 0; i < P.size(); i++) P[i] *= fs;
        for (size_t i = 0; i < Q.size(); i++) Q[i] *= fs;
    }
    for (int u = 0; u < m; u++) memcpy(P_out + (size_t)u * k, &P[(size_t)p_map[u] * k_al], sizeof(float) * k);
    for (int v = 0; v < n; v++) memcpy(Q_out + (size_t)v * k, &Q[(size_t)q_map[v] * k_al], sizeof(float) * k);
    *b_out = b;
    return 0;
}

// read_triplet, mf/mf.cpp:3367-3394: ids are truncated floats; m, n = max id + 1.
void orc_read_triplet(const float *tri, int count, orc_node *out, int *m, int *n) {
    int mm = 0, nn = 0;
    for (int j = 0; j < count; j++) {
        orc_node N;
        N.u = (int)tri[3 * j];
        N.v = (int)tri[3 * j + 1];
        N.r = tri[3 * j + 2];
        if (N.u + 1 > mm) mm = N.u + 1;
        if (N.v + 1 > nn) nn = N.v + 1;
        out[j] = N;
    }
    *m = mm;
    *n = nn;
}

// cos_similarity, mf/mf.cpp:3591-3683: the integer Q matrix from float triplets (read_triplet + 3610-3615, cells no
// triplet names are 0 here -- the reference leaves them uninitialised), cos = int dot / (sqrt(int) * sqrt(int)) in double
// rounded to float (3634-3650), then the exchange sort of 3652-3668 (strict <, so NaNs never move).  out[items] = ids.
int orc_cos_similarity(int item_id, const float *tri, int count, float *out, float *cos_out) {
    int items = 0, k = 0;
    for (int j = 0; j < count; j++) {
        const int u = (int)tri[3 * j], v = (int)tri[3 * j + 1];
        if (u + 1 > items) items = u + 1;
        if (v + 1 > k) k = v + 1;
    }
    if (!out) return items;
    std::vector<int> q((size_t)items * k, 0);
    for (int j = 0; j < count; j++) q[(size_t)(int)tri[3 * j] * k + (int)tri[3 * j + 1]] = (int)tri[3 * j + 2];
    std::vector<float> c((size_t)items), id((size_t)items);
    int item_abs = 0;
    for (int d = 0; d < k; d++) item_abs = item_abs + q[(size_t)item_id * k + d] * q[(size_t)item_id * k + d];
    for (int i = 0; i < items; i++) {
        int every_abs = 0, dot = 0;
        for (int d = 0; d < k; d++) {
            dot = dot + q[(size_t)item_id * k + d] * q[(size_t)i * k + d];
            every_abs = every_abs + q[(size_t)i * k + d] * q[(size_t)i * k + d];
        }
        c[(size_t)i] = dot / (std::sqrt((double)item_abs) * std::sqrt((double)every_abs));
        id[(size_t)i] = (float)i;
        if (cos_out) cos_out[i] = c[(size_t)i];
    }
    for (int i = 0; i < items - 1; i++)
        for (int j = i + 1; j < items; j++)
            if (c[(size_t)i] < c[(size_t)j]) {
                std::swap(c[(size_t)i], c[(size_t)j]);
                std::swap(id[(size_t)i], id[(size_t)j]);
            }
    for (int i = 0; i < items; i++) out[i] = id[(size_t)i];
    return items;
}

// utility_train, mf/mf.cpp:3483-3535 + model_to_array 3415-3441, in 1-thread order.
// `out` must hold 5 + m*k + n*k floats (caller sizes it via orc_read_triplet). Returns lens.
int orc_utility_train(const float *tri, int count, double p_l2, double q_l2, int k, int iters, double eta,
                      int rsqrt_mode, float *out) {
    std::vector<orc_node> R(count > 0 ? count : 1);
    int m, n;
    orc_read_triplet(tri, count, R.data(), &m, &n);
    orc_param prm;
    prm.k = k;
    prm.nr_bins = 20;  // mf_get_default_param, 4538-4557
    prm.nr_iters = iters;
    prm.lambda_p2 = (float)p_l2;
    prm.lambda_q2 = (float)q_l2;
    prm.eta = (float)eta;
    prm.rsqrt_mode = rsqrt_mode;
    float b;
    orc_train(R.data(), count, m, n, &prm, out + 5, out + 5 + (size_t)m * k, &b, nullptr, nullptr);
    out[0] = 0.f;  // P_L2_MFR
    out[1] = (float)m;
    out[2] = (float)n;
    out[3] = (float)k;
    out[4] = b;
    return 5 + m * k + n * k;
}

float orc_predict(const float *P, const float *Q, int m, int n, int k, float b, int u, int v) {
    return predict_one(P, Q, m, n, k, b, u, v);
}

// utility_predict, mf/mf.cpp:3537-3568: pairs are floats cast to int; one mf_predict per pair.
void orc_predict_pairs(const float *P, const float *Q, int m, int n, int k, float b, const float *pairs,
                       int npairs, float *out) {
    for (int i = 0; i < npairs; i++)
        out[i] = predict_one(P, Q, m, n, k, b, (int)pairs[2 * i], (int)pairs[2 * i + 1]);
}

// calc_rmse, mf/mf.cpp:4316-4331: float e, float e*e, double accumulation, sqrt(loss/nnz).
double orc_rmse(const orc_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k,
                float b) {
    if (nnz == 0) return 0;
    double loss = 0;
    for (long long i = 0; i < nnz; i++) {
        const float e = R[i].r - predict_one(P, Q, m, n, k, b, R[i].u, R[i].v);
        loss += e * e;
    }
    return std::sqrt(loss / nnz);
}

// calc_mae / calc_gkl / calc_logloss / calc_accuracy (mf/mf.cpp:4333-4404): `which` = 1 mae, 2 gkl, 5 logloss,
// 6 accuracy, anything else rmse.  Sums in double in index order (the reference's OpenMP reduction splits the
// range statically, so its sum can differ in the last bits from a sequential one: tests compare with 1e-12 rel).
double orc_metric(int which, const orc_node *R, long long nnz, const float *P, const float *Q, int m, int n,
                  int k, float b) {
    if (nnz == 0) return 0;
    double acc = 0;
    for (long long i = 0; i < nnz; i++) {
        const float z = predict_one(P, Q, m, n, k, b, R[i].u, R[i].v), r = R[i].r;
        switch (which) {
            case 1: acc += std::fabs(r - z); break;                       // float |r - z| widened
            case 2: acc += r * std::log(r / z) - r + z; break;            // all float, widened at the +=
            case 5: acc += r > 0 ? std::log(1.0 + std::exp(-z)) : std::log(1.0 + std::exp(z)); break;
            case 6: acc += r > 0 ? (z > 0 ? 1 : 0) : (z < 0 ? 1 : 0); break;
            default: {
                const float e = r - z;
                acc += (double)(e * e);
            }
        }
    }
    return which == 1 || which == 2 || which == 5 || which == 6 ? acc / nnz : std::sqrt(acc / nnz);
}

// calc_mpr_auc (mf/mf.cpp:4406-4525): for every row i with positives, all n scores by mf_predict; the positives' scores
// sorted ascending; for every non-positive column the number `left` of positives that do not beat it (binary search,
// 4480-4498): u_mpr += left, u_auc += pos - left; mpr = sum_i u_mpr / (n - pos) / total_pos, auc = sum_i u_auc / (n - pos)
// / pos / rows counted.  transpose: rows are the items.  Rows are summed in rising order (the reference's loop at one
// thread).  Ratings are taken as given: (row, column) pairs must be distinct (duplicates break the reference's swap loop).
void orc_mpr_auc(const orc_node *R, long long nnz, int prob_m, int prob_n, const float *P, const float *Q, int m, int n, int k,
                 float b, int transpose, double *out2) {
    const int rows = transpose ? std::max(prob_n, n) : std::max(prob_m, m);
    const int cols = transpose ? std::max(prob_m, m) : std::max(prob_n, n);
    std::vector<std::vector<int>> pos_of((size_t)rows);
    for (long long i = 0; i < nnz; i++) {
        const int r = transpose ? R[i].v : R[i].u, c = transpose ? R[i].u : R[i].v;
        if (R[i].r > 0) pos_of[(size_t)r].push_back(c);
    }
    int total_m = 0;
    long long total_pos = 0;
    double all_mpr = 0, all_auc = 0;
    std::vector<float> score((size_t)cols);
    std::vector<char> is_pos((size_t)cols);
    for (int i = 0; i < rows; i++) {
        const int pos = (int)pos_of[(size_t)i].size();
        if (pos < 1 || cols - pos < 1) continue;
        for (int j = 0; j < cols; j++)
            score[(size_t)j] = transpose ? predict_one(P, Q, m, n, k, b, j, i) : predict_one(P, Q, m, n, k, b, i, j);
        std::fill(is_pos.begin(), is_pos.end(), 0);
        std::vector<float> ps;
        for (int c : pos_of[(size_t)i]) {
            is_pos[(size_t)c] = 1;
            ps.push_back(score[(size_t)c]);
        }
        std::sort(ps.begin(), ps.end());
        total_m++;
        total_pos += pos;
        double u_mpr = 0, u_auc = 0;
        for (int j = 0; j < cols; j++) {
            if (is_pos[(size_t)j]) continue;
            const int left = (int)(std::upper_bound(ps.begin(), ps.end(), score[(size_t)j]) - ps.begin());
            u_mpr += left;
            u_auc += pos - left;
        }
        all_mpr += u_mpr / (cols - pos);
        all_auc += u_auc / (cols - pos) / pos;
    }
    out2[0] = all_mpr / (double)total_pos;
    out2[1] = all_auc / (double)total_m;
}


// Top-k oracle (no such function in the reference): score every item with
// mf_predict, order by (score desc, item id asc), keep the first `topk`.
void orc_topk(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers,
              int topk, int *idx_out, float *score_out) {
    std::vector<std::pair<float, int>> s(n);
    for (int i = 0; i < nusers; i++) {
        for (int v = 0; v < n; v++) s[v] = std::make_pair(predict_one(P, Q, m, n, k, b, users[i], v), v);
        const int kk = std::min(topk, n);
        std::partial_sort(s.begin(), s.begin() + kk, s.end(),
                          [](const std::pair<float, int> &a, const std::pair<float, int> &c) {
                              return a.first != c.first ? a.first > c.first : a.second < c.second;
                          });
        for (int j = 0; j < topk; j++) {
            idx_out[(size_t)i * topk + j] = j < kk ? s[j].second : -1;
            score_out[(size_t)i * topk + j] = j < kk ? s[j].first : 0.f;
        }
    }
}

// Synthetic ratings (counter based, so any slice can be produced anywhere).
static inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

void orc_gen_ratings(uint64_t seed, int m, int n, long long first, long long count, orc_node *out) {
#pragma omp parallel for schedule(static)  // counter based: every rating is a function of its index alone
    for (long long t = 0; t < count; t++) {
        const uint64_t i = (uint64_t)(first + t);
        const uint64_t h = sm64(seed ^ (i * 0x9E3779B97F4A7C15ull));
        const int u = (int)(h % (uint64_t)m), v = (int)((h >> 32) % (uint64_t)n);
        float z = 0.f;
        for (int d = 0; d < 8; d++) {
            const float a = u01(sm64(seed * 1000003ull + 1ull + 2ull * ((uint64_t)u * 8 + d))) * 0.9f;
            const float c = u01(sm64(seed * 1000003ull + 8ull + 2ull * ((uint64_t)v * 8 + d))) * 0.9f;
            z = z + a * c;
        }
        const float noise = u01(sm64(h)) - 0.5f;
        float r = 1.0f + z * 2.0f + noise;
        r = r < 1.f ? 1.f : (r > 5.f ? 5.f : r);
        out[t].u = u;
        out[t].v = v;
        out[t].r = r;
    }
}

// The same generator with skewed item popularity: P(item of popularity rank j) ~ 1/(j+1) (Zipf, exponent 1), built
// from integers only -- the octave of j+1 (its bit length) is uniform, j is uniform inside its octave.  At 17.8k items
// the most popular item receives 1/15 of all ratings.  Rank j is item id j; the engine's random permutation of the
// ids (gen_random_map, mf/mf.cpp:1009-1017) spreads the hot items over the grid.  Users stay uniform.
void orc_gen_ratings_zipf(uint64_t seed, int m, int n, long long first, long long count, orc_node *out) {
    int nb = 0;
    while (nb < 31 && ((1ll << nb) - 1) < (long long)n) nb++;  // octaves [2^o - 1, 2^(o+1) - 1) that start below n
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < count; t++) {
        const uint64_t i = (uint64_t)(first + t);
        const uint64_t h = sm64(seed ^ (i * 0x9E3779B97F4A7C15ull));
        const uint64_t g = sm64(h ^ 0xD1B54A32D192ED03ull);
        const int u = (int)(h % (uint64_t)m);
        const int o = (int)(g % (uint64_t)nb);
        const long long lo = (1ll << o) - 1, hi = std::min<long long>((1ll << (o + 1)) - 1, (long long)n);
        const int v = (int)(lo + (long long)((g >> 8) % (uint64_t)(hi - lo)));
        float z = 0.f;
        for (int d = 0; d < 8; d++) {
            const float a = u01(sm64(seed * 1000003ull + 1ull + 2ull * ((uint64_t)u * 8 + d))) * 0.9f;
            const float c = u01(sm64(seed * 1000003ull + 8ull + 2ull * ((uint64_t)v * 8 + d))) * 0.9f;
            z = z + a * c;
        }
        const float noise = u01(sm64(h)) - 0.5f;
        float r = 1.0f + z * 2.0f + noise;
        r = r < 1.f ? 1.f : (r > 5.f ? 5.f : r);
        out[t].u = u;
        out[t].v = v;
        out[t].r = r;
    }
}

// KAT helpers for tests (SURVEY.md Appendix B "Library-behaviour KATs").
void orc_kat_random_map(int size, int *out) {
    std::vector<int> a = random_map(size);
    memcpy(out, a.data(), sizeof(int) * size);
}
void orc_kat_minstd(int count, float *out) {
    Minstd g;
    for (int i = 0; i < count; i++) out[i] = g.u01();
}
void orc_kat_glibc_rand(unsigned seed, int count, int *out) {
    GlibcRand g(seed);
    for (int i = 0; i < count; i++) out[i] = g.next();
}
float orc_kat_rsqrt(float x, int mode) { return mode == 1 ? rsqrt12_hw(x) : rsqrt12_table(x); }

}  // extern "C"
