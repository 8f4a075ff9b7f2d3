// csrc/mf_api.cpp -- the drop-in surface: namespace mf (mangled C++ API of the reference's
// libmf.so, mf/mf.h:68-151), the php_* C-ABI of php_mf/mfWarp.h:6-10, and the mfb200_* C-ABI of
// include/mfb200.h.  Host glue only; compute goes through engine.cpp -> kernels.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <mutex>
#include <new>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/mf_b200.hpp"
#include "../../include/mfb200.h"
#include "engine.hpp"
#include "model_text.hpp"
#include "nccl_dl.hpp"

// layouts the reference's callers were compiled against (mf/mf.h:36-79)
static_assert(sizeof(mf::mf_node) == 12 && sizeof(mfb200_node) == 12 && sizeof(mfk_node) == 12, "mf_node layout");
static_assert(sizeof(mf::mf_problem) == 24, "mf_problem layout");
static_assert(sizeof(mf::mf_parameter) == 44, "mf_parameter layout");
static_assert(sizeof(mf::mf_model) == 40, "mf_model layout");

namespace {

std::mutex g_api_mutex;  // one GPU job at a time per process (PHP may be a ZTS build)

void not_supported(const char *what) {
    mfb200::set_error(std::string(what) + " is outside the accelerated path of this build (SURVEY.md section 2) "
                                          "and is not implemented");
}

float *aligned_floats(size_t count) {  // malloc_aligned_float, mf/mf.cpp:936-950: 32-byte aligned, free()-able
    void *p = nullptr;
    if (posix_memalign(&p, 32, sizeof(float) * (count ? count : 1)) != 0) throw std::bad_alloc();
    return (float *)p;
}

// check_parameter, mf/mf.cpp:3115-3184: same conditions, same messages on stderr.
bool params_ok(const mf::mf_parameter &p) {
    using namespace mf;
    auto fail = [](const char *msg) {
        std::cerr << msg << std::endl;
        return false;
    };
    const int f = p.fun;
    if (f != P_L2_MFR && f != P_L1_MFR && f != P_KL_MFR && f != P_LR_MFC && f != P_L2_MFC && f != P_L1_MFC &&
        f != P_ROW_BPR_MFOC && f != P_COL_BPR_MFOC)
        return fail("unknown loss function");
    if (p.k < 1) return fail("number of factors must be greater than zero");
    if (p.nr_threads < 1) return fail("number of threads must be greater than zero");
    if (p.nr_bins < 1 || p.nr_bins < p.nr_threads) return fail("number of bins must be greater than number of threads");
    if (p.nr_iters < 1) return fail("number of iterations must be greater than zero");
    if (p.lambda_p1 < 0 || p.lambda_p2 < 0 || p.lambda_q1 < 0 || p.lambda_q2 < 0)
        return fail("regularization coefficient must be non-negative");
    if (p.eta <= 0) return fail("learning rate must be greater than zero");
    if (f == P_KL_MFR && !p.do_nmf) return fail("--nmf must be set when using generalized KL-divergence");
    if (p.nr_bins <= 2 * p.nr_threads)
        std::cerr << "Warning: insufficient blocks may slow down the training"
                  << "process (4*nr_threads^2+1 blocks is suggested)" << std::endl;
    return true;
}

// MFB200_GPUS=N (N > 1): the one-shot training calls -- mf_train, utility_train, php_utility_train, mfb200_train -- use N
// devices of this host from inside the calling process: one host thread per device, each thread a rank of the S-stripe
// rotation (engine.cpp, world = N), the ranks joined by one NCCL communicator per device (kept for the process).  The
// caller sees the same call as on one GPU (mf/mf.h:89-91); rank 0 prints the table and hands the model over.
int train_multi(int ngpus, const mfb200_node *R, long long nnz, int m, int n, const mfb200_param &prm, float *P, float *Q,
                float *b, mfb200_report *rep, const mfb200_node *va, long long va_nnz) {
    const mfb200::NcclApi *nc = mfb200::nccl_api();
    if (!nc) return 1;
    ncclUniqueId id;
    if (nc->GetUniqueId(&id) != ncclSuccess) {
        mfb200::set_error("ncclGetUniqueId failed");
        return 1;
    }
    auto maps = std::make_shared<mfb200::SharedMaps>();
    std::vector<int> rcs((size_t)ngpus, 0);
    std::vector<std::string> errs((size_t)ngpus);
    mfb200_report r0;
    std::memset(&r0, 0, sizeof(r0));
    std::vector<std::thread> th;
    for (int r = 0; r < ngpus; r++)
        th.emplace_back([&, r] {
            mfb200_param pr = prm;
            pr.device = r;
            pr.mode = MFB200_MODE_RING;
            mfb200::Session s(m, n, pr, r, ngpus, &id);
            s.set_shared_maps(maps);
            s.set_silent(r != 0);
            s.set_validation(va, va_nnz);
            int rc = s.load(R, nnz) || s.run_epochs(prm.nr_iters, nullptr, nullptr, prm.quiet == 0) ||
                     s.finish(r == 0 ? P : nullptr, r == 0 ? Q : nullptr, r == 0 ? b : nullptr);
            if (rc) errs[(size_t)r] = mfb200::last_error();
            if (!rc && r == 0) s.fill_report(&r0);
            rcs[(size_t)r] = rc;
            s.release();
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < ngpus; r++)
        if (rcs[(size_t)r]) {
            mfb200::set_error("rank " + std::to_string(r) + " of " + std::to_string(ngpus) + ": " + errs[(size_t)r]);
            return 1;
        }
    r0.gpus = ngpus;
    if (rep) *rep = r0;
    return 0;
}

int gpus_wanted(long long nnz, const mfb200_param &prm) {
    const char *e = std::getenv("MFB200_GPUS");
    const int want = e && *e ? std::atoi(e) : 1;
    if (want <= 1) return 1;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) return 1;
    // several GPUs need the throughput mode (the exact mode replays ONE sequential order)
    const char *x = std::getenv("MFB200_EXACT_MAX_NNZ");
    const long long exact_max = x && *x ? std::atoll(x) : 262144;
    const bool ring = prm.mode == MFB200_MODE_RING || prm.mode == MFB200_MODE_RING_REPRO ||
                      (prm.mode == MFB200_MODE_AUTO && nnz > exact_max);
    if (!ring || ((prm.k + 7) / 8) * 8 > 512 || prm.fun == MFK_FUN_ROW_BPR || prm.fun == MFK_FUN_COL_BPR) return 1;
    return std::min(want, have);
}

int train_impl(const mfb200_node *R, long long nnz, int m, int n, const mfb200_param &prm, float *P, float *Q,
               float *b, mfb200_report *rep, const mfb200_node *va = nullptr, long long va_nnz = 0) {
    typedef std::chrono::steady_clock clk;
    auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
    const auto t0 = clk::now();
    mfb200_report r;
    std::memset(&r, 0, sizeof(r));
    int rc = 0;
    const int ngpus = gpus_wanted(nnz, prm);
    if (ngpus > 1) {
        rc = train_multi(ngpus, R, nnz, m, n, prm, P, Q, b, &r, va, va_nnz);
    } else {
        mfb200::Session s(m, n, prm);
        s.set_validation(va, va_nnz);
        rc = s.load(R, nnz) || s.run_epochs(prm.nr_iters, nullptr, nullptr, prm.quiet == 0) || s.finish(P, Q, b);
        if (!rc) s.fill_report(&r);
        const auto t1 = clk::now();
        s.release();  // what the destructor does, inside the clock: every phase of the call is accounted for
        r.destroy_ms = ms_since(t1);
    }
    r.total_ms = ms_since(t0);
    if (rep && !rc) *rep = r;
    return rc ? 1 : 0;
}

struct DevBuf {  // from the device's memory pool: a warm call pays no cudaMalloc / cudaFree
    void *p = nullptr;
    ~DevBuf() { mfb200::api_pool_free(p); }
    int alloc(size_t bytes) { return mfb200::api_pool_alloc(&p, bytes); }
};

int need_device() {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess || c == 0) {
        mfb200::set_error("no CUDA device available: this build has no CPU fallback");
        return 1;
    }
    const char *d = std::getenv("MFB200_DEVICE");  // the one-shot calls run on this device (default: current)
    if (d && *d && cudaSetDevice(std::atoi(d)) != cudaSuccess) {
        mfb200::set_error("MFB200_DEVICE names an unusable device");
        return 1;
    }
    return 0;
}

}  // namespace

// =================================================================================================
// C ABI (include/mfb200.h)
// =================================================================================================
extern "C" {

int mfb200_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}
const char *mfb200_last_error(void) { return mfb200::last_error(); }
const char *mfb200_version(void) { return "mfb200 0.1 (sm_100a)"; }

mfb200_param mfb200_default_param(void) {  // mf_get_default_param, mf/mf.cpp:4538-4557
    mfb200_param p;
    p.k = 8;
    p.nr_bins = 20;
    p.nr_iters = 20;
    p.lambda_p2 = 0.1f;
    p.lambda_q2 = 0.1f;
    p.eta = 0.1f;
    p.quiet = 0;
    p.mode = MFB200_MODE_AUTO;
    p.device = -1;
    p.fun = 0;
    p.lambda_p1 = 0.0f;
    p.lambda_q1 = 0.0f;
    p.do_nmf = 0;
    return p;
}

int mfb200_train(const mfb200_node *R, long long nnz, int m, int n, const mfb200_param *param, float *P_out,
                 float *Q_out, float *b_out, mfb200_report *report) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (!param || param->k < 1 || param->nr_iters < 1 || param->eta <= 0 || param->lambda_p2 < 0 || param->lambda_q2 < 0) {
        mfb200::set_error("invalid parameter");
        return 1;
    }
    return train_impl(R, nnz, m, n, *param, P_out, Q_out, b_out, report);
}

// ---- a model resident on the device (utility_predict re-parses and re-copies the whole model
// on every call, mf/mf.cpp:3559; VERDICT r1: P and Q were uploaded by every predict / metric / top-k call) ---------
}  // extern "C"
struct mfb200_model {
    int device = 0, m = 0, n = 0, k = 0;
    float b = 0.f;
    float *dP = nullptr, *dQ = nullptr;
};
namespace {
thread_local double t_topk_ms = 0.0, t_eval_ms = 0.0;

struct EventTimer {  // device time of what is enqueued on the legacy stream between start() and stop()
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    EventTimer() {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
    }
    ~EventTimer() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    void start() { cudaEventRecord(e0, nullptr); }
    double stop() {
        cudaEventRecord(e1, nullptr);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms;
    }
};

// uploads P and Q into pooled device memory of the current device
int model_upload(mfb200_model &M, const float *P, const float *Q, int m, int n, int k, float b) {
    if (m < 0 || n < 0 || k < 1 || (m > 0 && !P) || (n > 0 && !Q)) {
        mfb200::set_error("invalid model");
        return 1;
    }
    if (need_device()) return 1;
    cudaGetDevice(&M.device);
    M.m = m; M.n = n; M.k = k; M.b = b;
    if (mfb200::api_pool_alloc((void **)&M.dP, sizeof(float) * (size_t)m * k) ||
        mfb200::api_pool_alloc((void **)&M.dQ, sizeof(float) * (size_t)n * k)) {
        mfb200::set_error("cudaMalloc failed");
        return 1;
    }
    return mfb200::api_h2d(M.dP, P, sizeof(float) * (size_t)m * k) || mfb200::api_h2d(M.dQ, Q, sizeof(float) * (size_t)n * k);
}
void model_release(mfb200_model &M) {
    mfb200::api_pool_free(M.dP);
    mfb200::api_pool_free(M.dQ);
    M.dP = M.dQ = nullptr;
}
struct ScopedModel {  // the one-shot calls: a model that lives for one call
    mfb200_model M;
    ~ScopedModel() { model_release(M); }
};

// utility_predict's loop over mf_predict (mf/mf.cpp:3562-3565) on a resident model
int predict_core(const mfb200_model &M, const float *pairs, long long npairs, float *out) {
    if (npairs <= 0) return 0;
    DevBuf dpairs, dout;
    if (dpairs.alloc(sizeof(float) * 2 * (size_t)npairs) || dout.alloc(sizeof(float) * (size_t)npairs)) {
        mfb200::set_error("cudaMalloc failed");
        return 1;
    }
    if (mfb200::api_h2d(dpairs.p, pairs, sizeof(float) * 2 * (size_t)npairs)) return 1;
    EventTimer tm;
    tm.start();
    int rc = mfk_predict_pairs(M.dP, M.dQ, M.m, M.n, M.k, M.b, (const float *)dpairs.p, npairs, (float *)dout.p, nullptr);
    t_eval_ms = tm.stop();
    if (!rc && mfb200::api_d2h(out, dout.p, sizeof(float) * (size_t)npairs)) return 1;
    if (rc) mfb200::set_error(std::string("predict failed: ") + cudaGetErrorString((cudaError_t)rc));
    return rc ? 1 : 0;
}

// calc_rmse / calc_mae / calc_gkl / calc_logloss / calc_accuracy (mf/mf.cpp:4316-4404) on a resident model
int metric_core(int which, const mfb200_model &M, const mfb200_node *R, long long nnz, double *out) {
    if (nnz == 0) {  // mf/mf.cpp:4318-4319
        *out = 0;
        return 0;
    }
    DevBuf dR, dacc;
    if (dR.alloc(sizeof(mfb200_node) * (size_t)nnz) || dacc.alloc(sizeof(double))) {
        mfb200::set_error("cudaMalloc failed");
        return 1;
    }
    if (mfb200::api_h2d(dR.p, R, sizeof(mfb200_node) * (size_t)nnz)) return 1;
    cudaMemsetAsync(dacc.p, 0, sizeof(double), nullptr);
    EventTimer tm;
    tm.start();
    int rc;
    if (which == MFK_FUN_L2_MFR)
        rc = mfk_sq_err((const mfk_node *)dR.p, nnz, M.dP, M.dQ, M.m, M.n, M.k, M.b, (double *)dacc.p, nullptr);
    else
        rc = mfk_err_general(which, (const mfk_node *)dR.p, nnz, nullptr, nullptr, M.dP, M.dQ, M.m, M.n, M.k, M.b, 1.0f,
                             (double *)dacc.p, 0, mfk_hidden{nullptr, 1, 1, 1}, nullptr);
    t_eval_ms = tm.stop();
    double loss = 0;
    if (!rc) rc = (int)cudaMemcpy(&loss, dacc.p, sizeof(double), cudaMemcpyDeviceToHost);
    if (rc) {
        mfb200::set_error(std::string("metric failed: ") + cudaGetErrorString((cudaError_t)rc));
        return 1;
    }
    *out = which == MFK_FUN_L2_MFR ? std::sqrt(loss / (double)nnz) : loss / (double)nnz;
    return 0;
}

// batched top-k on a resident model.  Users whose candidate list overflows in the GEMM path (many items tied at the cut)
// and shapes that path does not take (k > 128 or topk > 128 with more than 2048 items) go through the exact path, one
// user at a time: every item scored like mf_predict, full sort.
int topk_core(const mfb200_model &M, const int *users, int nusers, int topk, int *idx_out, float *score_out) {
    if (nusers <= 0) return 0;
    if (!users || !idx_out || M.n < 1 || topk < 1) {
        mfb200::set_error("mfb200_topk: invalid argument");
        return 1;
    }
    const int m = M.m, n = M.n, k = M.k;
    const int sms = mfk_sm_count(M.device);
    const int batch = std::min(((nusers + 255) / 256) * 256, std::max(1, sms) * 256);  // users per GEMM batch
    const int n_tiles = (n + 255) / 256;
    const char *se = std::getenv("MFB200_TOPK_STRIDE");
    int stride = se && *se ? std::atoi(se) : (n_tiles >= 16 * topk ? 2 : 1);
    stride = std::max(1, std::min(stride, 8));
    const bool gemm_shape = n <= 2048 ? topk <= 2048 : (((k + 63) / 64) * 64 <= 128 && topk <= 128);
    const size_t wbytes = !gemm_shape ? 256 : n > 2048 ? mfk_topk_work_bytes(n, k, batch, stride) : 256;
    DevBuf dU, dI, dS, dW, dO;
    if (dU.alloc(sizeof(int) * (size_t)nusers) || dI.alloc(sizeof(int) * (size_t)nusers * topk) ||
        dS.alloc(sizeof(float) * (size_t)nusers * topk) || dW.alloc(wbytes) || dO.alloc(sizeof(int) * ((size_t)nusers + 1))) {
        mfb200::set_error("mfb200_topk: cudaMalloc failed");
        return 1;
    }
    if (mfb200::api_h2d(dU.p, users, sizeof(int) * (size_t)nusers)) return 1;
    cudaMemsetAsync(dO.p, 0, sizeof(int), nullptr);
    EventTimer tm;
    tm.start();
    int rc = 0, overflow = 0;
    std::vector<int> redo;
    if (gemm_shape) {
        rc = mfk_topk(M.dP, M.dQ, m, n, k, M.b, (const int *)dU.p, nusers, topk, (int *)dI.p, (float *)dS.p, dW.p, wbytes, batch,
                      stride, sms, (int *)dO.p, nullptr);
        if (!rc) rc = (int)cudaMemcpy(&overflow, dO.p, sizeof(int), cudaMemcpyDeviceToHost);
        if (!rc && overflow > 0) {
            redo.resize((size_t)std::min(overflow, nusers));
            rc = (int)cudaMemcpy(redo.data(), (const int *)dO.p + 1, sizeof(int) * redo.size(), cudaMemcpyDeviceToHost);
        }
    } else {
        redo.resize((size_t)nusers);
        for (int i = 0; i < nusers; i++) redo[(size_t)i] = i;
    }
    if (!rc && !redo.empty()) {
        DevBuf dX;
        const size_t xbytes = mfk_topk_exact_work_bytes(n);
        if (dX.alloc(xbytes)) {
            mfb200::set_error("mfb200_topk: cudaMalloc failed");
            return 1;
        }
        for (size_t i = 0; i < redo.size() && !rc; i++)
            rc = mfk_topk_exact_user(M.dP, M.dQ, m, n, k, M.b, (const int *)dU.p, redo[i], topk, (int *)dI.p, (float *)dS.p, dX.p,
                                     xbytes, nullptr);
        cudaStreamSynchronize(nullptr);  // dX goes back to the pool
    }
    t_topk_ms = tm.stop();  // device time of the scoring itself (factors already resident)
    if (!rc && mfb200::api_d2h(idx_out, dI.p, sizeof(int) * (size_t)nusers * topk)) return 1;
    if (!rc && score_out && mfb200::api_d2h(score_out, dS.p, sizeof(float) * (size_t)nusers * topk)) return 1;
    if (rc) {
        mfb200::set_error(std::string("mfb200_topk failed: ") + cudaGetErrorString((cudaError_t)rc));
        return 1;
    }
    return 0;
}

int metric_impl(int which, const mfb200_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k, float b,
                double *out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (nnz == 0) {
        *out = 0;
        return 0;
    }
    ScopedModel sm;
    if (model_upload(sm.M, P, Q, m, n, k, b)) return 1;
    return metric_core(which, sm.M, R, nnz, out);
}
}  // namespace
extern "C" {

mfb200_model *mfb200_model_upload(const float *P, const float *Q, int m, int n, int k, float b) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    mfb200_model *M = new (std::nothrow) mfb200_model();
    if (!M) return nullptr;
    if (model_upload(*M, P, Q, m, n, k, b)) {
        model_release(*M);
        delete M;
        return nullptr;
    }
    cudaStreamSynchronize(nullptr);  // the caller may free or change P and Q as soon as the call returns
    return M;
}
void mfb200_model_free(mfb200_model *M) {
    if (!M) return;
    std::lock_guard<std::mutex> lock(g_api_mutex);
    cudaSetDevice(M->device);
    model_release(*M);
    delete M;
}
int mfb200_model_predict_pairs(const mfb200_model *M, const float *pairs, long long npairs, float *out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (!M || cudaSetDevice(M->device) != cudaSuccess) {
        mfb200::set_error("invalid model handle");
        return 1;
    }
    return predict_core(*M, pairs, npairs, out);
}
int mfb200_model_metric(const mfb200_model *M, int which, const mfb200_node *R, long long nnz, double *out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (!M || !out || cudaSetDevice(M->device) != cudaSuccess) {
        mfb200::set_error("invalid model handle");
        return 1;
    }
    if (which != MFK_FUN_L2_MFR && which != MFK_FUN_L1_MFR && which != MFK_FUN_KL_MFR && which != MFK_FUN_LR_MFC &&
        which != MFK_FUN_L2_MFC && which != MFK_FUN_L1_MFC) {
        mfb200::set_error("mfb200_model_metric: unknown error measure");
        return 1;
    }
    return metric_core(which, *M, R, nnz, out);
}
int mfb200_model_topk(const mfb200_model *M, const int *users, int nusers, int topk, int *idx_out, float *score_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (!M || cudaSetDevice(M->device) != cudaSuccess) {
        mfb200::set_error("invalid model handle");
        return 1;
    }
    return topk_core(*M, users, nusers, topk, idx_out, score_out);
}
double mfb200_eval_last_ms(void) { return t_eval_ms; }

int mfb200_predict_pairs(const float *P, const float *Q, int m, int n, int k, float b, const float *pairs,
                         long long npairs, float *out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (npairs <= 0) return 0;
    ScopedModel sm;
    if (model_upload(sm.M, P, Q, m, n, k, b)) return 1;
    return predict_core(sm.M, pairs, npairs, out);
}

int mfb200_rmse(const mfb200_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k, float b,
                double *rmse_out) {
    return metric_impl(MFK_FUN_L2_MFR, R, nnz, P, Q, m, n, k, b, rmse_out);
}

int mfb200_metric(int which, const mfb200_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k,
                  float b, double *out) {
    if (which != MFK_FUN_L2_MFR && which != MFK_FUN_L1_MFR && which != MFK_FUN_KL_MFR && which != MFK_FUN_LR_MFC &&
        which != MFK_FUN_L2_MFC && which != MFK_FUN_L1_MFC) {
        mfb200::set_error("mfb200_metric: unknown error measure");
        return 1;
    }
    return metric_impl(which, R, nnz, P, Q, m, n, k, b, out);
}

// mf_cross_validation (mf/mf.cpp:4117-4129, CrossValidatorBase::do_cross_validation 3208-3262): the blocks of the
// nr_bins x nr_bins grid are shuffled (srand(0), random_shuffle) and dealt to the folds; every fold trains with its
// blocks hidden and takes the loss's error measure over them.  In exact mode a fold is the reference's fold bit for bit
// (an epoch is nr_bins^2 jobs over the blocks that are left).  The throughput schedule makes whole passes over the
// training ratings instead, so it runs round(nr_iters * nblk / (nblk - hidden)) of them -- the same number of visits.
int mfb200_cross_validation(const mfb200_node *R, long long nnz, int m, int n, const mfb200_param *param, int nr_folds,
                            double *fold_errors, double *mean_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (!param || !mean_out || nr_folds < 1) {
        mfb200::set_error("mfb200_cross_validation: invalid argument");
        return 1;
    }
    const int bins = std::max(1, param->nr_bins), nblk = bins * bins, bpf = nblk / nr_folds;
    if (bpf < 1 || bpf >= nblk) {
        mfb200::set_error("mfb200_cross_validation: the number of folds must be between 2 and nr_bins^2");
        return 1;
    }
    const std::vector<int> order = mfb200::cv_block_order(nblk);
    double sum = 0;
    for (int f = 0; f < nr_folds; f++) {
        const int lo = f * bpf, hi = std::min((f + 1) * bpf, nblk);
        mfb200_param prm = *param;
        prm.quiet = 1;  // CrossValidatorBase's constructor silences the folds (mf/mf.cpp:3210)
        mfb200::Session s(m, n, prm);
        s.set_hidden_blocks(order.data() + lo, hi - lo);
        if (s.load(R, nnz)) return 1;
        int epochs = prm.nr_iters;
        if (s.mode_used() == MFB200_MODE_RING)
            epochs = (int)std::lround((double)prm.nr_iters * nblk / (double)(nblk - (hi - lo)));
        if (s.run_epochs(epochs, nullptr, nullptr, false)) return 1;
        double err = 0;
        if (s.cv_error(&err)) return 1;
        if (fold_errors) fold_errors[f] = err;
        sum += err;
    }
    *mean_out = sum / nr_folds;
    return 0;
}

double mfb200_topk_last_ms(void) { return t_topk_ms; }

int mfb200_topk(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers, int topk,
                int *idx_out, float *score_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (nusers <= 0) return 0;
    ScopedModel sm;
    if (model_upload(sm.M, P, Q, m, n, k, b)) return 1;
    return topk_core(sm.M, users, nusers, topk, idx_out, score_out);
}

// SURVEY.md 8d generator; product-side copy (the oracle has its own, tests compare the two).
static inline unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline float unit24(unsigned long long h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

void mfb200_gen_ratings(unsigned long long seed, int m, int n, long long first, long long count, mfb200_node *out) {
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < count; t++) {
        const unsigned long long i = (unsigned long long)(first + t);
        const unsigned long long h = mix64(seed ^ (i * 0x9E3779B97F4A7C15ull));
        const int u = (int)(h % (unsigned long long)m), v = (int)((h >> 32) % (unsigned long long)n);
        float z = 0.f;
        for (int d = 0; d < 8; d++) {
            const float pu = unit24(mix64(seed * 1000003ull + 1ull + 2ull * ((unsigned long long)u * 8 + d))) * 0.9f;
            const float qv = unit24(mix64(seed * 1000003ull + 8ull + 2ull * ((unsigned long long)v * 8 + d))) * 0.9f;
            z = z + pu * qv;
        }
        float r = 1.0f + z * 2.0f + (unit24(mix64(h)) - 0.5f);
        out[t].u = u;
        out[t].v = v;
        out[t].r = r < 1.f ? 1.f : (r > 5.f ? 5.f : r);
    }
}

// ---- staged sessions ----------------------------------------------------------------------------
struct mfb200_session {
    mfb200::Session impl;
    mfb200_session(int m, int n, const mfb200_param &p, int rank = 0, int world = 1, const void *id = nullptr)
        : impl(m, n, p, rank, world, id) {}
};

// ---- calc_mpr_auc (mf/mf.cpp:4406-4525): the ranking measures of the one-class losses ----
int mfb200_mpr_auc(const mfb200_node *R, long long nnz, int prob_m, int prob_n, const float *P, const float *Q, int m, int n,
                   int k, float b, int transpose, double *mpr_out, double *auc_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (!R || nnz <= 0 || !mpr_out || !auc_out) {
        mfb200::set_error("mfb200_mpr_auc: invalid argument");
        return 1;
    }
    ScopedModel sm;
    if (model_upload(sm.M, P, Q, m, n, k, b)) return 1;
    const int rows = transpose ? std::max(prob_n, n) : std::max(prob_m, m);
    const int cols = transpose ? std::max(prob_m, m) : std::max(prob_n, n);
    const size_t tmp_bytes = mfk_rank_sort_tmp_bytes(nnz);
    // rows per batch: the scores of a batch (rows x columns floats) stay below 512 MB
    int batch = (int)std::max<long long>(1, std::min<long long>(rows, (128ll << 20) / std::max(cols, 1)));
    if (const char *e = std::getenv("MFB200_RANK_BATCH_ROWS"))
        if (std::atoi(e) > 0) batch = std::min(batch, std::atoi(e));
    DevBuf dR, dK0, dK1, dKept, dStart, dScores, dPos, dPosSorted, dMpr, dAuc, dTmp;
    if (dR.alloc(sizeof(mfb200_node) * (size_t)nnz) || dK0.alloc(8 * (size_t)nnz) || dK1.alloc(8 * (size_t)nnz) ||
        dKept.alloc(8) || dStart.alloc(8 * ((size_t)rows + 1)) || dScores.alloc(4 * (size_t)batch * cols) ||
        dPos.alloc(4 * (size_t)nnz) || dPosSorted.alloc(4 * (size_t)nnz) || dMpr.alloc(8 * (size_t)rows) ||
        dAuc.alloc(8 * (size_t)rows) || dTmp.alloc(tmp_bytes)) {
        mfb200::set_error("cudaMalloc failed");
        return 1;
    }
    if (mfb200::api_h2d(dR.p, R, sizeof(mfb200_node) * (size_t)nnz)) return 1;
    int rc = mfk_rank_prepare((const mfk_node *)dR.p, nnz, transpose, rows, (unsigned long long *)dK0.p,
                              (unsigned long long *)dK1.p, (unsigned long long *)dKept.p, (long long *)dStart.p, dTmp.p,
                              tmp_bytes, nullptr);
    std::vector<long long> start((size_t)rows + 1, 0);
    if (!rc) rc = (int)cudaMemcpy(start.data(), dStart.p, 8 * start.size(), cudaMemcpyDeviceToHost);
    for (int lo = 0; lo < rows && !rc; lo += batch) {
        const int hi = std::min(rows, lo + batch);
        rc = mfk_rank_batch(sm.M.dP, sm.M.dQ, m, n, k, b, transpose, lo, hi, cols, (const unsigned long long *)dK1.p,
                            (const long long *)dStart.p, start[(size_t)lo], start[(size_t)hi], (float *)dScores.p,
                            (float *)dPos.p, (float *)dPosSorted.p, (unsigned long long *)dMpr.p,
                            (unsigned long long *)dAuc.p, dTmp.p, tmp_bytes, nullptr);
    }
    std::vector<unsigned long long> um((size_t)rows), ua((size_t)rows);
    if (!rc) rc = (int)cudaMemcpy(um.data(), dMpr.p, 8 * um.size(), cudaMemcpyDeviceToHost);
    if (!rc) rc = (int)cudaMemcpy(ua.data(), dAuc.p, 8 * ua.size(), cudaMemcpyDeviceToHost);
    if (rc) {
        mfb200::set_error(std::string("mfb200_mpr_auc failed: ") + cudaGetErrorString((cudaError_t)rc));
        return 1;
    }
    // the reference's sums, row by row in rising order (4448-4513 at one thread)
    int total_m = 0;
    long long total_pos = 0;
    double all_mpr = 0, all_auc = 0;
    for (int i = 0; i < rows; i++) {
        const long long pos = start[(size_t)i + 1] - start[(size_t)i];
        if (pos < 1 || cols - pos < 1) continue;
        total_m++;
        total_pos += pos;
        all_mpr += (double)um[(size_t)i] / (double)(cols - pos);
        all_auc += (double)ua[(size_t)i] / (double)(cols - pos) / (double)pos;
    }
    *mpr_out = all_mpr / (double)total_pos;
    *auc_out = all_auc / (double)total_m;
    return 0;
}

// ---- cosine similarity of Q-matrix rows (mf::cos_similarity, mf/mf.cpp:3591-3683) ----
namespace {
// read_triplet (mf/mf.cpp:3367-3394: float -> int truncation, sizes = largest index + 1) and the fill loop of
// cos_similarity (3600-3615: q_array[u][v] = (int) r).  Cells no triplet names are 0 here; the reference leaves them
// uninitialised (malloc) -- a caller that relies on that gets whatever the heap held.
bool q_matrix_from_triplets(const float *tri, int count, int &items, int &k, std::vector<int> &q) {
    items = 0;
    k = 0;
    if (!tri || count <= 0) return false;
    for (int j = 0; j < count; j++) {
        const int u = (int)tri[3 * j], v = (int)tri[3 * j + 1];
        if (u < 0 || v < 0) return false;
        items = std::max(items, u + 1);
        k = std::max(k, v + 1);
    }
    if ((long long)items * k > (1ll << 31)) return false;
    q.assign((size_t)items * k, 0);
    for (int j = 0; j < count; j++) q[(size_t)(int)tri[3 * j] * k + (int)tri[3 * j + 1]] = (int)tri[3 * j + 2];
    return true;
}
}  // namespace

int mfb200_cos_similarity(const float *q_triplets, int n_triplets, const int *item_ids, int n_ids, int *order_out,
                          float *cos_sorted_out, float *cos_by_item_out, int *ties_out, int *items_out, int *k_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    int items = 0, k = 0;
    std::vector<int> q;
    if (!q_matrix_from_triplets(q_triplets, n_triplets, items, k, q)) {
        mfb200::set_error("cos_similarity: invalid Q-matrix triplets");
        return 1;
    }
    if (items_out) *items_out = items;
    if (k_out) *k_out = k;
    if (!order_out && !cos_sorted_out && !cos_by_item_out) return 0;  // size query
    if (n_ids <= 0 || !item_ids) n_ids = items, item_ids = nullptr;  // all items
    for (int i = 0; item_ids && i < n_ids; i++)
        if (item_ids[i] < 0 || item_ids[i] >= items) {
            mfb200::set_error("cos_similarity: item id out of range");
            return 1;
        }
    if (need_device()) return 1;
    const size_t total = (size_t)n_ids * items, tmp_bytes = mfk_cos_tmp_bytes(items, n_ids);
    DevBuf dQ, dA, dC, dI, dCs, dIs, dF, dT;
    if (dQ.alloc(sizeof(int) * q.size()) || dA.alloc(sizeof(int) * (size_t)n_ids) || dC.alloc(sizeof(float) * total) ||
        dI.alloc(sizeof(int) * total) || dCs.alloc(sizeof(float) * total) || dIs.alloc(sizeof(int) * total) ||
        dF.alloc(sizeof(int) * (size_t)n_ids) || dT.alloc(tmp_bytes)) {
        mfb200::set_error("cudaMalloc failed");
        return 1;
    }
    if (mfb200::api_h2d(dQ.p, q.data(), sizeof(int) * q.size())) return 1;
    if (item_ids && mfb200::api_h2d(dA.p, item_ids, sizeof(int) * (size_t)n_ids)) return 1;
    cudaError_t e = cudaMemsetAsync(dF.p, 0, sizeof(int) * (size_t)n_ids, nullptr);
    int rc = e != cudaSuccess ? (int)e
                              : mfk_cos_similarity((const int *)dQ.p, items, k, item_ids ? (const int *)dA.p : nullptr, n_ids,
                                                   (float *)dC.p, (int *)dI.p, (float *)dCs.p, (int *)dIs.p, (int *)dF.p, dT.p,
                                                   tmp_bytes, nullptr);
    if (rc) {
        mfb200::set_error(std::string("cos_similarity kernels: ") + cudaGetErrorString((cudaError_t)rc));
        return 1;
    }
    if ((order_out && mfb200::api_d2h(order_out, dIs.p, sizeof(int) * total)) ||
        (cos_sorted_out && mfb200::api_d2h(cos_sorted_out, dCs.p, sizeof(float) * total)) ||
        (cos_by_item_out && mfb200::api_d2h(cos_by_item_out, dC.p, sizeof(float) * total)) ||
        (ties_out && mfb200::api_d2h(ties_out, dF.p, sizeof(int) * (size_t)n_ids)))
        return 1;
    return 0;
}

int mfb200_dist_unique_id(unsigned char id128[128]) {
    const mfb200::NcclApi *nc = mfb200::nccl_api();
    if (!nc || !id128) return 1;
    ncclUniqueId id;
    if (nc->GetUniqueId(&id) != ncclSuccess) {
        mfb200::set_error("ncclGetUniqueId failed");
        return 1;
    }
    std::memcpy(id128, &id, 128);
    return 0;
}

mfb200_session *mfb200_dist_session_create(int m, int n, const mfb200_param *param, int rank, int world,
                                           const unsigned char id128[128]) {
    if (!param || param->k < 1 || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) {
        mfb200::set_error("invalid parameter");
        return nullptr;
    }
    return new (std::nothrow) mfb200_session(m, n, *param, rank, world, id128);
}

long long mfb200_dist_exchange_plan(int world, int rank, const unsigned long long *counts, long long *send_off,
                                    long long *recv_off) {
    if (world < 1 || rank < 0 || rank >= world || !counts || !send_off || !recv_off) return -1;
    return mfb200::exchange_plan(world, rank, counts, send_off, recv_off);
}
int mfb200_dist_owner_of_row(int t_row, int t_seg, int world) { return mfb200::owner_of_row(t_row, t_seg, world); }

void mfb200_dist_rotation(int world, int rank, long long substep, int stripes_per_rank, int out5[5]) {
    const mfb200::RotationStep r = mfb200::rotation_step(world, rank, substep, stripes_per_rank < 1 ? 1 : stripes_per_rank);
    out5[0] = r.compute;
    out5[1] = r.send_stripe;
    out5[2] = r.send_to;
    out5[3] = r.recv_stripe;
    out5[4] = r.recv_from;
}

namespace {
// the plan of the default loss with locks, as Session::load picks it
bool plan_default(int m, int n, long long nnz, int k, int world, int rank, int sm_count, int max_smem, mfk_band_shape *s) {
    const char *kn = std::getenv("MFB200_KERNEL");
    int kind = ((k + 7) / 8) * 8 <= 128 ? 4 : 0;
    if (kn && !std::strcmp(kn, "band")) kind = 0;
    if (kn && !std::strcmp(kn, "run") && kind) kind = 1;
    if (kn && !std::strcmp(kn, "cell") && kind) kind = 2;
    if (kn && !std::strcmp(kn, "warp") && kind) kind = 3;
    if (kn && !std::strcmp(kn, "tlock") && kind) kind = 5;
    if (kn && !std::strcmp(kn, "item") && kind) kind = 6;
    return mfb200::plan_band(m, n, nnz, ((k + 7) / 8) * 8, sm_count, max_smem, world, rank, s, kind);
}
}  // namespace

// which SGD kernel that plan launches: the codes of mfb200_report.kernel (1 band, 2 run, 3 cell, 4 warp, 5 run with T-row
// locks, 6 item); -1 on failure
int mfb200_plan_kernel(int m, int n, long long nnz, int k, int world, int rank, int sm_count, int max_smem) {
    mfk_band_shape s;
    if (!plan_default(m, n, nnz, k, world, rank, sm_count, max_smem, &s)) return -1;
    return s.by_row == 4 ? 6 : s.by_row == 3 ? 4 : s.by_row == 2 ? 3 : s.by_row ? (s.tlock ? 5 : 2) : 1;
}

int mfb200_plan_band(int m, int n, long long nnz, int k, int world, int rank, int sm_count, int max_smem,
                     int out16[16]) {
    mfk_band_shape s;
    if (!plan_default(m, n, nnz, k, world, rank, sm_count, max_smem, &s)) return 1;
    const int v[16] = {s.nC, s.nWarps, s.L, s.nG, s.S1, s.nTB, s.nPass, s.segS, s.segT, s.segT2, s.swap_sides,
                       s.nStripes, s.stripeRows, s.tLo, s.tRows, (int)s.smem_bytes};
    std::memcpy(out16, v, sizeof(v));
    return 0;
}

mfb200_session *mfb200_session_create(int m, int n, const mfb200_param *param) {
    if (!param || param->k < 1) {
        mfb200::set_error("invalid parameter");
        return nullptr;
    }
    return new (std::nothrow) mfb200_session(m, n, *param);
}
int mfb200_session_load(mfb200_session *s, const mfb200_node *R_host, long long nnz) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    return s ? s->impl.load(R_host, nnz) : 1;
}
int mfb200_session_set_validation(mfb200_session *s, const mfb200_node *va_host, long long nnz) {
    if (!s) return 1;
    s->impl.set_validation(va_host, nnz);
    return 0;
}
int mfb200_session_reset(mfb200_session *s) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    return s ? s->impl.reset() : 1;
}
int mfb200_session_epochs(mfb200_session *s, int epochs, float *ms_out, double *tr_rmse_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    return s ? s->impl.run_epochs(epochs, ms_out, tr_rmse_out, false) : 1;
}
int mfb200_session_finish(mfb200_session *s, float *P_out, float *Q_out, float *b_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    return s ? s->impl.finish(P_out, Q_out, b_out) : 1;
}
int mfb200_session_rmse(mfb200_session *s, const mfb200_node *R_host, long long nnz, double *rmse_out) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    return s ? s->impl.heldout_rmse(R_host, nnz, rmse_out) : 1;
}
int mfb200_session_report(mfb200_session *s, mfb200_report *report) {
    if (!s || !report) return 1;
    s->impl.fill_report(report);
    return 0;
}
void *mfb200_session_stream(mfb200_session *s) { return s ? s->impl.stream() : nullptr; }
void mfb200_session_destroy(mfb200_session *s) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    delete s;
}

// ---- the reference's own C-ABI (php_mf/mfWarp.cpp:3-34) -------------------------------------------
int php_mf_my_train(char *tr_path, char *model_path) { return mf::mf_my_train(tr_path, model_path); }
float *php_utility_train(float *train_data, int train_triplet_num, double p_l2, double q_l2, int k, int iters,
                         double eta, int *lens) {
    return mf::utility_train(train_data, train_triplet_num, p_l2, q_l2, k, iters, eta, *lens);
}
float *php_utility_predict(float *test_arr, int test_triplet_num, float *model_arr, int model_arr_len) {
    return mf::utility_predict(test_arr, test_triplet_num, model_arr, model_arr_len);
}
float *php_cos_similarity(int item_id, float *q_arr, int q_arr_num) { return mf::cos_similarity(item_id, q_arr, q_arr_num); }
int *php_DINA(float *q_arr, int q_triplet_num, float *x_arr, int x_triplet_num, int iterators) {
    return mf::DINA(q_arr, q_triplet_num, x_arr, x_triplet_num, iterators);
}

}  // extern "C"

// =================================================================================================
// namespace mf: the mangled C++ surface
// =================================================================================================
namespace mf {

mf_parameter mf_get_default_param() {  // mf/mf.cpp:4538-4557
    mf_parameter p;
    std::memset(&p, 0, sizeof(p));
    p.fun = P_L2_MFR;
    p.k = 8;
    p.nr_threads = 12;
    p.nr_bins = 20;
    p.nr_iters = 20;
    p.lambda_p1 = 0.0f;
    p.lambda_q1 = 0.0f;
    p.lambda_p2 = 0.1f;
    p.lambda_q2 = 0.1f;
    p.eta = 0.1f;
    p.do_nmf = false;
    p.quiet = false;
    p.copy_data = true;
    return p;
}

mf_model *mf_train_with_validation(mf_problem const *tr, mf_problem const *va, mf_parameter param) {
    if (!params_ok(param)) return nullptr;  // mf/mf.cpp:3312-3313
    if (!tr) {
        mfb200::set_error("null training problem");
        return nullptr;
    }
    mfb200_param prm = mfb200_default_param();
    prm.k = param.k;
    prm.nr_bins = param.nr_bins;
    prm.nr_iters = param.nr_iters;
    prm.lambda_p2 = param.lambda_p2;
    prm.lambda_q2 = param.lambda_q2;
    prm.eta = param.eta;
    prm.quiet = param.quiet ? 1 : 0;
    prm.fun = param.fun;
    prm.lambda_p1 = param.lambda_p1;
    prm.lambda_q1 = param.lambda_q1;
    prm.do_nmf = param.do_nmf ? 1 : 0;
    const char *mode = std::getenv("MFB200_MODE");  // mf_parameter cannot grow: the mode comes from the environment
    if (mode && !std::strcmp(mode, "exact")) prm.mode = MFB200_MODE_EXACT;
    if (mode && !std::strcmp(mode, "ring")) prm.mode = MFB200_MODE_RING;
    if (mode && !std::strcmp(mode, "ring_repro")) prm.mode = MFB200_MODE_RING_REPRO;
    // the one-class BPR losses (mf/mf.cpp:2131-2707) exist in the exact mode only, whatever the environment says
    if (param.fun == P_ROW_BPR_MFOC || param.fun == P_COL_BPR_MFOC) prm.mode = MFB200_MODE_EXACT;

    mf_model *model = new mf_model;
    model->fun = param.fun;
    model->m = tr->m;
    model->n = tr->n;
    model->k = param.k;
    model->b = 0;
    model->P = model->Q = nullptr;
    try {
        model->P = aligned_floats((size_t)tr->m * param.k);
        model->Q = aligned_floats((size_t)tr->n * param.k);
    } catch (std::bad_alloc const &e) {
        std::cerr << e.what() << std::endl;  // mf/mf.cpp:980-985
        mf_destroy_model(&model);
        throw;
    }
    int rc;
    {
        std::lock_guard<std::mutex> lock(g_api_mutex);
        rc = train_impl((const mfb200_node *)tr->R, tr->nnz, tr->m, tr->n, prm, model->P, model->Q, &model->b, nullptr,
                        va ? (const mfb200_node *)va->R : nullptr, va ? va->nnz : 0);
    }
    if (rc) {
        mf_destroy_model(&model);
        return nullptr;
    }
    return model;
}

mf_model *mf_train(mf_problem const *prob, mf_parameter param) { return mf_train_with_validation(prob, nullptr, param); }

void mf_destroy_model(mf_model **model) {  // mf/mf.cpp:4280-4293
    if (!model || !*model) return;
    std::free((*model)->P);
    std::free((*model)->Q);
    delete *model;
    *model = nullptr;
}

// A single pair is a host scalar by nature (a kernel launch per pair would cost ~10^4 times the
// arithmetic); batches go through utility_predict / calc_rmse, which run on the GPU.
mf_float mf_predict(mf_model const *model, mf_int u, mf_int v) {  // mf/mf.cpp:4295-4314
    if (u < 0 || u >= model->m || v < 0 || v >= model->n) return model->b;
    const mf_float *p = model->P + (mf_long)u * model->k, *q = model->Q + (mf_long)v * model->k;
    mf_float z = 0.0f;
    for (mf_int d = 0; d < model->k; d++) z = z + p[d] * q[d];
    return std::isnan(z) ? model->b : z;
}

mf_double calc_rmse(mf_problem *prob, mf_model *model) {  // mf/mf.cpp:4316-4331
    double out = std::numeric_limits<double>::quiet_NaN();
    if (mfb200_rmse((const mfb200_node *)prob->R, prob->nnz, model->P, model->Q, model->m, model->n, model->k,
                    model->b, &out))
        return std::numeric_limits<double>::quiet_NaN();
    return out;
}

mf_problem read_triplet(float *tri, int triplet_num) {  // mf/mf.cpp:3367-3394
    mf_problem prob;
    prob.m = prob.n = 0;
    prob.nnz = triplet_num;
    prob.R = new mf_node[triplet_num > 0 ? triplet_num : 1];
    for (int j = 0; j < triplet_num; j++) {
        mf_node &N = prob.R[j];
        N.u = (mf_int)tri[3 * j];      // float -> int truncation, as the reference
        N.v = (mf_int)tri[3 * j + 1];
        N.r = tri[3 * j + 2];
        if (N.u + 1 > prob.m) prob.m = N.u + 1;
        if (N.v + 1 > prob.n) prob.n = N.v + 1;
    }
    return prob;
}

float *model_to_array(mf_model *model, int &lens) {  // mf/mf.cpp:3415-3441
    const size_t pn = (size_t)model->m * model->k, qn = (size_t)model->n * model->k;
    lens = (int)(pn + qn + 5);
    float *out = (float *)std::malloc(sizeof(float) * (pn + qn + 5));
    if (!out) return nullptr;
    out[0] = (float)model->fun;
    out[1] = (float)model->m;
    out[2] = (float)model->n;
    out[3] = (float)model->k;
    out[4] = model->b;
    std::memcpy(out + 5, model->P, sizeof(float) * pn);
    std::memcpy(out + 5 + pn, model->Q, sizeof(float) * qn);
    return out;
}

mf_model *array_to_model(float *arr, int lens) {  // mf/mf.cpp:3444-3481
    if (!arr || lens < 5) return nullptr;
    mf_model *model = new mf_model;
    model->fun = (mf_int)arr[0];
    model->m = (mf_int)arr[1];
    model->n = (mf_int)arr[2];
    model->k = (mf_int)arr[3];
    model->b = arr[4];
    model->P = model->Q = nullptr;
    const long long pn = (long long)model->m * model->k, qn = (long long)model->n * model->k;
    if (model->m < 0 || model->n < 0 || model->k < 0 || (long long)lens != pn + qn + 5) {
        delete model;
        return nullptr;
    }
    model->P = (float *)std::malloc(sizeof(float) * (size_t)(pn ? pn : 1));
    model->Q = (float *)std::malloc(sizeof(float) * (size_t)(qn ? qn : 1));
    std::memcpy(model->P, arr + 5, sizeof(float) * (size_t)pn);
    std::memcpy(model->Q, arr + 5 + pn, sizeof(float) * (size_t)qn);
    return model;
}

float *utility_train(float *train_data, int train_triplet_num, double p_l2, double q_l2, int k, int iters, double eta,
                     int &lens) {  // mf/mf.cpp:3483-3535
    mf_problem tr = read_triplet(train_data, train_triplet_num);
    mf_parameter param = mf_get_default_param();
    param.lambda_p2 = (mf_float)p_l2;
    param.lambda_q2 = (mf_float)q_l2;
    param.k = k;
    param.nr_iters = iters;
    param.eta = (mf_float)eta;
    mf_model *model = mf_train_with_validation(&tr, nullptr, param);
    delete[] tr.R;
    if (!model) {  // the reference would dereference the null model here
        lens = 0;
        return nullptr;
    }
    float *arr = model_to_array(model, lens);
    mf_destroy_model(&model);  // the reference leaks the model (mf/mf.cpp:3530); nothing observes that
    return arr;
}

float *utility_predict(float *test_arr, int test_triplet_num, float *model_arr, int model_arr_len) {  // mf/mf.cpp:3537-3568
    mf_model *model = array_to_model(model_arr, model_arr_len);
    if (!model) {  // length mismatch: the reference crashes in mf_predict; fail loudly instead
        mfb200::set_error("utility_predict: model array length does not match its header");
        return nullptr;
    }
    float *out = (float *)std::malloc(sizeof(float) * (size_t)(test_triplet_num > 0 ? test_triplet_num : 1));
    int rc = mfb200_predict_pairs(model->P, model->Q, model->m, model->n, model->k, model->b, test_arr,
                                  test_triplet_num, out);
    mf_destroy_model(&model);
    if (rc) {
        std::free(out);
        return nullptr;
    }
    return out;
}

// ---- text formats (host) ---------------------------------------------------------------------------
mf_problem read_problem(char const *path) {  // mf/mf.cpp:4143-4182: lines "u v r"
    mf_problem prob;
    prob.m = prob.n = 0;
    prob.nnz = 0;
    prob.R = nullptr;
    if (!path) return prob;
    static_assert(sizeof(mfb200::TextNode) == sizeof(mf_node), "TextNode must have the layout of mf_node");
    mfb200::TextNode *nodes = nullptr;
    long long nnz = 0;
    // (the caller releases R with delete[], like the reference's: mf/mf.cpp:4161)
    if (mfb200::read_problem_text(path, [](unsigned long long c) { return (mfb200::TextNode *)new mf_node[c]; }, &nodes,
                                  &nnz, &prob.m, &prob.n))
        return prob;
    prob.nnz = nnz;
    prob.R = (mf_node *)nodes;
    return prob;
}

mf_int mf_save_model(mf_model const *model, char const *path) {  // mf/mf.cpp:4184-4225
    if (!model) return 1;
    return mfb200::save_model_text(path, model->fun, model->m, model->n, model->k, model->b, model->P, model->Q);
}

mf_model *mf_load_model(char const *path) {  // mf/mf.cpp:4227-4278
    mfb200::ModelTextHeader h;
    float *P = nullptr, *Q = nullptr;
    try {
        if (mfb200::load_model_text(path, &h, [](unsigned long long c) { return aligned_floats((size_t)c); }, &P, &Q)) {
            std::free(P);
            std::free(Q);
            return nullptr;
        }
    } catch (std::bad_alloc const &e) {
        std::cerr << e.what() << std::endl;  // mf/mf.cpp:4248-4253
        std::free(P);
        std::free(Q);
        return nullptr;
    }
    mf_model *model = new mf_model;
    model->fun = h.fun;
    model->m = h.m;
    model->n = h.n;
    model->k = h.k;
    model->b = h.b;
    model->P = P;
    model->Q = Q;
    return model;
}

mf_int mf_my_train(char const *tr_path, char const *model_path) {  // mf/mf.cpp:3397-3413
    mf_problem tr = read_problem(tr_path);
    mf_parameter param = mf_get_default_param();
    param.nr_iters = 40;
    mf_model *model = mf_train_with_validation(&tr, nullptr, param);
    mf_int status = model ? mf_save_model(model, model_path) : -1;
    mf_destroy_model(&model);
    delete[] tr.R;
    return status;
}

// ---- outside the path: exported so dependants link; fail loudly --------------------------------------
// cos_similarity (mf/mf.cpp:3591-3683): item ids (as floats) by falling cosine with `item_id`.  The cosines and a sorted
// list come from the device (csrc/cos_sim.cu).  With distinct cosines every correct sort returns that list.  With equal
// cosines or NaNs (zero rows) the reference's answer is whatever its exchange sort (3652-3668: for i, for j > i, swap when
// cos[i] < cos[j]) leaves behind; that sequence of swaps is then replayed on the device's cosines so that the list is the
// reference's, entry for entry.  The result is malloc'd and owned by the caller (who, in PHP, never frees it).
// On bad input the reference reads out of bounds; here a message goes to stderr and the list is all zeros -- never
// NULL, because php_mf.c:1211 dereferences the result without a check.
float *cos_similarity(int item_id, float *q_arr, int q_arr_num) {
    int items = 0, k = 0;
    if (mfb200_cos_similarity(q_arr, q_arr_num, nullptr, 0, nullptr, nullptr, nullptr, nullptr, &items, &k) || item_id < 0 ||
        item_id >= items) {
        mfb200::set_error("cos_similarity: invalid Q-matrix triplets or item id");
        return (float *)std::calloc((size_t)std::max(items, 1), sizeof(float));
    }
    std::vector<int> order((size_t)items);
    std::vector<float> by_item((size_t)items);
    int ties = 0;
    float *result = (float *)std::calloc((size_t)items, sizeof(float));
    if (!result) return nullptr;
    if (mfb200_cos_similarity(q_arr, q_arr_num, &item_id, 1, order.data(), nullptr, by_item.data(), &ties, nullptr, nullptr))
        return result;  // (message already on stderr)
    if (!ties) {
        for (int i = 0; i < items; i++) result[i] = (float)order[(size_t)i];
        return result;
    }
    std::vector<float> id((size_t)items);
    for (int i = 0; i < items; i++) id[(size_t)i] = (float)i;
    float *c = by_item.data();
    for (int i = 0; i + 1 < items; i++)
        for (int j = i + 1; j < items; j++)
            if (c[i] < c[j]) {
                std::swap(c[i], c[j]);
                std::swap(id[(size_t)i], id[(size_t)j]);
            }
    std::memcpy(result, id.data(), sizeof(float) * (size_t)items);
    return result;
}
// DINA (mf/mf.cpp:3685-4115) is outside the accelerated path (out of scope).  It is exported so
// that libphp_mf.so links.  php_mf.c:1281-1287 reads 20 ints of the result without a check, so the call reports the
// condition on stderr and returns a zeroed, malloc'd buffer of that size instead of NULL.
int *DINA(float *, int, float *, int, int) {
    not_supported("DINA");
    return (int *)std::calloc(64, sizeof(int));
}
mf_model *mf_train_on_disk(char const *, mf_parameter) {
    not_supported("mf_train_on_disk");
    return nullptr;
}
mf_model *mf_train_with_validation_on_disk(char const *, char const *, mf_parameter) {
    not_supported("mf_train_with_validation_on_disk");
    return nullptr;
}
mf_double mf_cross_validation(mf_problem const *prob, mf_int nr_folds, mf_parameter param) {  // mf/mf.cpp:4117-4129
    if (!params_ok(param)) return 0;
    if (param.fun == P_ROW_BPR_MFOC || param.fun == P_COL_BPR_MFOC) {
        not_supported("one-class (BPR) cross-validation");
        return std::numeric_limits<double>::quiet_NaN();
    }
    if (!prob || nr_folds < 1) return std::numeric_limits<double>::quiet_NaN();
    mfb200_param prm = mfb200_default_param();
    prm.k = param.k;
    prm.nr_bins = param.nr_bins;
    prm.nr_iters = param.nr_iters;
    prm.lambda_p2 = param.lambda_p2;
    prm.lambda_q2 = param.lambda_q2;
    prm.eta = param.eta;
    prm.fun = param.fun;
    prm.lambda_p1 = param.lambda_p1;
    prm.lambda_q1 = param.lambda_q1;
    prm.do_nmf = param.do_nmf ? 1 : 0;
    const char *mode = std::getenv("MFB200_MODE");
    if (mode && !std::strcmp(mode, "exact")) prm.mode = MFB200_MODE_EXACT;
    if (mode && !std::strcmp(mode, "ring")) prm.mode = MFB200_MODE_RING;
    if (mode && !std::strcmp(mode, "ring_repro")) prm.mode = MFB200_MODE_RING_REPRO;
    std::vector<double> errs((size_t)nr_folds, 0.0);
    double mean = 0;
    if (mfb200_cross_validation((const mfb200_node *)prob->R, prob->nnz, prob->m, prob->n, &prm, nr_folds, errs.data(),
                                &mean))
        return std::numeric_limits<double>::quiet_NaN();
    if (!param.quiet) {  // the table of do_cross_validation, mf/mf.cpp:3216-3259
        static const char *legend[] = {"rmse", "mae", "gkl", "", "", "logloss", "accuracy", "accuracy"};
        std::cout.width(4);
        std::cout << "fold";
        std::cout.width(10);
        std::cout << legend[param.fun & 7];
        std::cout << std::endl;
        for (int f = 0; f < nr_folds; f++) {
            std::cout.width(4);
            std::cout << f;
            std::cout.width(10);
            std::cout << std::fixed << std::setprecision(4) << errs[(size_t)f];
            std::cout << std::endl;
        }
        std::cout.width(14);
        std::cout.fill('=');
        std::cout << "" << std::endl;
        std::cout.fill(' ');
        std::cout.width(4);
        std::cout << "avg";
        std::cout.width(10);
        std::cout << std::fixed << std::setprecision(4) << mean;
        std::cout << std::endl;
    }
    return mean;
}
mf_double mf_cross_validation_on_disk(char const *, mf_int, mf_parameter) {
    not_supported("mf_cross_validation_on_disk");
    return std::numeric_limits<double>::quiet_NaN();
}
#define MFB200_METRIC(name, which)                                                                                 \
    mf_double name(mf_problem *prob, mf_model *model) {                                                            \
        double out = std::numeric_limits<double>::quiet_NaN();                                                     \
        if (mfb200_metric(which, (const mfb200_node *)prob->R, prob->nnz, model->P, model->Q, model->m, model->n,  \
                          model->k, model->b, &out))                                                               \
            return std::numeric_limits<double>::quiet_NaN();                                                       \
        return out;                                                                                                \
    }
MFB200_METRIC(calc_mae, MFK_FUN_L1_MFR)       // mf/mf.cpp:4333-4347
MFB200_METRIC(calc_gkl, MFK_FUN_KL_MFR)       // mf/mf.cpp:4349-4364
MFB200_METRIC(calc_logloss, MFK_FUN_LR_MFC)   // mf/mf.cpp:4366-4384
MFB200_METRIC(calc_accuracy, MFK_FUN_L2_MFC)  // mf/mf.cpp:4386-4404
// calc_mpr / calc_auc (mf/mf.cpp:4406-4536) on the device (csrc/rank_metrics.cu).  Like the reference they sort the
// caller's rating array by (row, column) -- (column, row) when transposed -- as a side effect (4432).
static std::pair<mf_double, mf_double> mpr_auc_impl(mf_problem *prob, mf_model *model, bool transpose) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (!prob || !model) return std::make_pair(nan, nan);
    if (transpose)
        std::sort(prob->R, prob->R + prob->nnz, [](const mf_node &a, const mf_node &b) { return std::tie(a.v, a.u) < std::tie(b.v, b.u); });
    else
        std::sort(prob->R, prob->R + prob->nnz, [](const mf_node &a, const mf_node &b) { return std::tie(a.u, a.v) < std::tie(b.u, b.v); });
    double mpr = nan, auc = nan;
    if (mfb200_mpr_auc((const mfb200_node *)prob->R, prob->nnz, prob->m, prob->n, model->P, model->Q, model->m, model->n,
                       model->k, model->b, transpose ? 1 : 0, &mpr, &auc))
        return std::make_pair(nan, nan);
    return std::make_pair(mpr, auc);
}
std::pair<mf_double, mf_double> calc_mpr_auc(mf_problem *prob, mf_model *model, bool transpose) {
    return mpr_auc_impl(prob, model, transpose);
}
mf_double calc_mpr(mf_problem *prob, mf_model *model, bool transpose) { return mpr_auc_impl(prob, model, transpose).first; }
mf_double calc_auc(mf_problem *prob, mf_model *model, bool transpose) { return mpr_auc_impl(prob, model, transpose).second; }

}  // namespace mf
