// oracle/ref_shim.cpp
//
// TEST INFRASTRUCTURE ONLY.  A C-ABI doorway into the UNMODIFIED reference library
// (oracle/_ref/libmf_ref.so, compiled by oracle/Makefile from /root/reference/mf/mf.cpp
// where it lies).  It exists so that tests/ and bench.py can drive the reference with ctypes:
// the reference's own API is C++ (namespace mf, mf_parameter passed by value, mf/mf.h:19-23).
// Nothing here computes anything; every function forwards to the reference.
//
// Built only where /root/reference exists (this container); the resulting
// oracle/_ref/libref_shim.so travels to the GPU box as a prebuilt file.

#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <streambuf>
#include <vector>

#include "/root/reference/mf/mf.h"

namespace {

// Swallows the reference's per-iteration table (mf/mf.cpp:2880-2907) and records the time
// at which each line ended: line 0 is the header, line i+1 ends epoch i.
class StampBuf : public std::streambuf {
public:
    std::vector<double> stamps;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();

protected:
    int overflow(int c) override {
        if (c == '\n')
            stamps.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return c;
    }
    std::streamsize xsputn(const char *s, std::streamsize n) override {
        for (std::streamsize i = 0; i < n; i++) overflow(s[i]);
        return n;
    }
};

}  // namespace

extern "C" {

// mf_train (mf/mf.cpp:3362-3365) with every parameter of the path exposed.
// stamps (may be null) receives up to max_stamps line-end times in seconds since the call
// started; *n_stamps the number written.  total_s = wall time of mf_train itself.
int ref_train(const mf::mf_node *R, long long nnz, int m, int n, int k, int nr_bins, int nr_iters,
              int nr_threads, float lambda_p2, float lambda_q2, float eta, float *P_out, float *Q_out,
              float *b_out, double *stamps, int max_stamps, int *n_stamps, double *total_s) {
    mf::mf_problem prob;
    prob.m = m;
    prob.n = n;
    prob.nnz = nnz;
    prob.R = const_cast<mf::mf_node *>(R);
    mf::mf_parameter prm = mf::mf_get_default_param();
    prm.k = k;
    prm.nr_bins = nr_bins;
    prm.nr_iters = nr_iters;
    prm.nr_threads = nr_threads;
    prm.lambda_p2 = lambda_p2;
    prm.lambda_q2 = lambda_q2;
    prm.eta = eta;
    prm.quiet = (stamps == nullptr);

    StampBuf sb;
    std::streambuf *old = std::cout.rdbuf(&sb);
    const std::ios::fmtflags oldf = std::cout.flags();
    auto t0 = std::chrono::steady_clock::now();
    sb.t0 = t0;
    mf::mf_model *mdl = mf::mf_train(&prob, prm);
    auto t1 = std::chrono::steady_clock::now();
    std::cout.rdbuf(old);
    std::cout.flags(oldf);
    if (total_s) *total_s = std::chrono::duration<double>(t1 - t0).count();
    if (n_stamps) {
        int c = (int)sb.stamps.size() < max_stamps ? (int)sb.stamps.size() : max_stamps;
        for (int i = 0; i < c; i++) stamps[i] = sb.stamps[i];
        *n_stamps = c;
    }
    if (!mdl) return 1;
    if (P_out) memcpy(P_out, mdl->P, sizeof(float) * (size_t)mdl->m * mdl->k);
    if (Q_out) memcpy(Q_out, mdl->Q, sizeof(float) * (size_t)mdl->n * mdl->k);
    if (b_out) *b_out = mdl->b;
    mf::mf_destroy_model(&mdl);
    return 0;
}

// mf_train with the remaining knobs of mf_parameter (loss function, L1 regularisation, NMF).  The per-iteration
// table the reference prints (mf/mf.cpp:2818-2907) is returned as text in `table` (up to table_cap-1 bytes, NUL
// terminated) so that its tr_<metric> and obj columns can be compared.
int ref_train_ex(const mf::mf_node *R, long long nnz, int m, int n, int k, int nr_bins, int nr_iters,
                 int nr_threads, int fun, float lambda_p1, float lambda_q1, float lambda_p2, float lambda_q2,
                 float eta, int do_nmf, float *P_out, float *Q_out, float *b_out, char *table, int table_cap) {
    mf::mf_problem prob;
    prob.m = m;
    prob.n = n;
    prob.nnz = nnz;
    prob.R = const_cast<mf::mf_node *>(R);
    mf::mf_parameter prm = mf::mf_get_default_param();
    prm.fun = fun;
    prm.k = k;
    prm.nr_bins = nr_bins;
    prm.nr_iters = nr_iters;
    prm.nr_threads = nr_threads;
    prm.lambda_p1 = lambda_p1;
    prm.lambda_q1 = lambda_q1;
    prm.lambda_p2 = lambda_p2;
    prm.lambda_q2 = lambda_q2;
    prm.eta = eta;
    prm.do_nmf = do_nmf != 0;
    prm.quiet = (table == nullptr);

    std::stringbuf sb;
    std::streambuf *old = std::cout.rdbuf(&sb);
    const std::ios::fmtflags oldf = std::cout.flags();
    mf::mf_model *mdl = mf::mf_train(&prob, prm);
    std::cout.rdbuf(old);
    std::cout.flags(oldf);
    if (table && table_cap > 0) {
        const std::string t = sb.str();
        const size_t c = t.size() < (size_t)table_cap - 1 ? t.size() : (size_t)table_cap - 1;
        memcpy(table, t.data(), c);
        table[c] = 0;
    }
    if (!mdl) return 1;
    if (P_out) memcpy(P_out, mdl->P, sizeof(float) * (size_t)mdl->m * mdl->k);
    if (Q_out) memcpy(Q_out, mdl->Q, sizeof(float) * (size_t)mdl->n * mdl->k);
    if (b_out) *b_out = mdl->b;
    mf::mf_destroy_model(&mdl);
    return 0;
}

// mf_cross_validation (mf/mf.cpp:4117-4129) with every parameter of the path exposed; quiet.
double ref_cross_validation(const mf::mf_node *R, long long nnz, int m, int n, int k, int nr_bins, int nr_iters,
                            int nr_threads, int fun, float lambda_p1, float lambda_q1, float lambda_p2,
                            float lambda_q2, float eta, int do_nmf, int nr_folds) {
    mf::mf_problem prob;
    prob.m = m;
    prob.n = n;
    prob.nnz = nnz;
    prob.R = const_cast<mf::mf_node *>(R);
    mf::mf_parameter prm = mf::mf_get_default_param();
    prm.fun = fun;
    prm.k = k;
    prm.nr_bins = nr_bins;
    prm.nr_iters = nr_iters;
    prm.nr_threads = nr_threads;
    prm.lambda_p1 = lambda_p1;
    prm.lambda_q1 = lambda_q1;
    prm.lambda_p2 = lambda_p2;
    prm.lambda_q2 = lambda_q2;
    prm.eta = eta;
    prm.do_nmf = do_nmf != 0;
    prm.quiet = true;
    const std::ios::fmtflags oldf = std::cout.flags();
    const double r = mf::mf_cross_validation(&prob, nr_folds, prm);
    std::cout.flags(oldf);
    return r;
}

// The other metrics of mf.h (mf/mf.cpp:4333-4404) on a caller-provided model: 1 mae, 2 gkl, 5 logloss, 6 accuracy.
double ref_metric(int which, const mf::mf_node *R, long long nnz, const float *P, const float *Q, int m, int n,
                  int k, float b) {
    mf::mf_problem prob;
    prob.m = m;
    prob.n = n;
    prob.nnz = nnz;
    prob.R = const_cast<mf::mf_node *>(R);
    mf::mf_model mdl;
    mdl.fun = 0;
    mdl.m = m;
    mdl.n = n;
    mdl.k = k;
    mdl.b = b;
    mdl.P = const_cast<float *>(P);
    mdl.Q = const_cast<float *>(Q);
    switch (which) {
        case 1: return mf::calc_mae(&prob, &mdl);
        case 2: return mf::calc_gkl(&prob, &mdl);
        case 5: return mf::calc_logloss(&prob, &mdl);
        case 6: return mf::calc_accuracy(&prob, &mdl);
        default: return mf::calc_rmse(&prob, &mdl);
    }
}

float *ref_utility_train(float *tri, int count, double p_l2, double q_l2, int k, int iters, double eta,
                         int *lens) {
    // NOTE: trains with nr_threads=12 (mf/mf.cpp:4544): not reproducible, and on tiny inputs
    // it can dead-lock after the last epoch -- callers use a watchdog.
    std::streambuf *old = std::cout.rdbuf(nullptr);
    const std::ios::fmtflags oldf = std::cout.flags();
    float *out = mf::utility_train(tri, count, p_l2, q_l2, k, iters, eta, *lens);
    std::cout.rdbuf(old);
    std::cout.flags(oldf);
    return out;
}

float *ref_utility_predict(float *pairs, int npairs, float *model_arr, int model_len) {
    return mf::utility_predict(pairs, npairs, model_arr, model_len);
}

float ref_predict(const float *P, const float *Q, int m, int n, int k, float b, int u, int v) {
    mf::mf_model mdl;
    mdl.fun = 0;
    mdl.m = m;
    mdl.n = n;
    mdl.k = k;
    mdl.b = b;
    mdl.P = const_cast<float *>(P);
    mdl.Q = const_cast<float *>(Q);
    return mf::mf_predict(&mdl, u, v);
}

double ref_rmse(const mf::mf_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k,
                float b) {
    mf::mf_problem prob;
    prob.m = m;
    prob.n = n;
    prob.nnz = nnz;
    prob.R = const_cast<mf::mf_node *>(R);
    mf::mf_model mdl;
    mdl.fun = 0;
    mdl.m = m;
    mdl.n = n;
    mdl.k = k;
    mdl.b = b;
    mdl.P = const_cast<float *>(P);
    mdl.Q = const_cast<float *>(Q);
    return mf::calc_rmse(&prob, &mdl);
}

// cos_similarity, mf/mf.cpp:3591-3683 (the caller passes a triplet for EVERY cell: the reference reads the cells no
// triplet names uninitialised)
float *ref_cos_similarity(int item_id, float *q_arr, int q_arr_num) { return mf::cos_similarity(item_id, q_arr, q_arr_num); }

// calc_mpr_auc (mf/mf.cpp:4406-4525) through calc_mpr / calc_auc; the reference sorts prob->R in place: a copy is passed
void ref_mpr_auc(const mf::mf_node *R, long long nnz, int prob_m, int prob_n, const float *P, const float *Q, int m, int n,
                 int k, float b, int transpose, double *out2) {
    std::vector<mf::mf_node> copy(R, R + nnz);
    mf::mf_problem prob;
    prob.m = prob_m;
    prob.n = prob_n;
    prob.nnz = nnz;
    prob.R = copy.data();
    mf::mf_model mdl;
    mdl.fun = 10;
    mdl.m = m;
    mdl.n = n;
    mdl.k = k;
    mdl.b = b;
    mdl.P = const_cast<float *>(P);
    mdl.Q = const_cast<float *>(Q);
    out2[0] = mf::calc_mpr(&prob, &mdl, transpose != 0);
    out2[1] = mf::calc_auc(&prob, &mdl, transpose != 0);
}

void ref_free(void *p) { free(p); }

// the process-wide rand() state the reference's Scheduler seeds its per-block generators from (mf/mf.cpp:103-110)
void ref_srand(unsigned seed) { srand(seed); }

}  // extern "C"
