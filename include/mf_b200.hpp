// include/mf_b200.hpp -- the C++ drop-in surface of libmf.so (B200 build).
//
// The reference's library API is C++ in namespace mf with Itanium-mangled names (its
// `extern "C"` is commented out, mf/mf.h:19-23), and mf_parameter travels BY VALUE
// (mf/mf.h:89-91).  A drop-in libmf.so therefore has to export exactly those mangled symbols with
// byte-identical POD layouts.  Callers compiled against the reference's own mf/mf.h (php_mf/
// mfWarp.cpp, mfTest/mfTest.cpp) bind to the symbols declared here without recompilation of
// their sources; this header exists for new C++ callers and for the layout static_asserts in
// csrc/mf_api.cpp.  Each declaration cites what it replaces.
//
// Scope: the train / predict / metric path of the matrix-factorisation solver runs on
// the GPU for the six MFSolver losses (fun = P_L2_MFR, the one the PHP surface uses, and P_L1_MFR, P_KL_MFR,
// P_LR_MFC, P_L2_MFC, P_L1_MFC) with L1/L2 regularisation and NMF, cross-validation included; the two one-class BPR
// losses (P_ROW_BPR_MFOC, P_COL_BPR_MFOC) train in the exact mode, with calc_mpr / calc_auc; cos_similarity runs on the
// device.  Entry points outside that path (on-disk training, BPR cross-validation, DINA) are exported so that dependants
// link, and fail loudly (message on stderr; null / NaN / non-zero result, a zeroed buffer where the PHP glue would
// dereference null).
#ifndef MF_B200_HPP
#define MF_B200_HPP

#define MFB200_EXPORT __attribute__((visibility("default")))

namespace mf {

typedef float mf_float;      // mf/mf.h:26-29
typedef double mf_double;
typedef int mf_int;
typedef long long mf_long;

// loss ids, mf/mf.h:31-32.  The two one-class (BPR) losses train in the exact mode only.
enum { P_L2_MFR = 0, P_L1_MFR = 1, P_KL_MFR = 2, P_LR_MFC = 5, P_L2_MFC = 6, P_L1_MFC = 7,
       P_ROW_BPR_MFOC = 10, P_COL_BPR_MFOC = 11 };

struct mf_node {             // mf/mf.h:36-41, 12 bytes
    mf_int u, v;
    mf_float r;
};

struct mf_problem {          // mf/mf.h:43-49, 24 bytes
    mf_int m, n;
    mf_long nnz;
    mf_node *R;
};

struct mf_parameter {        // mf/mf.h:51-66, 44 bytes: 5 int, 5 float, 3 bool (+1 pad)
    mf_int fun, k, nr_threads, nr_bins, nr_iters;
    mf_float lambda_p1, lambda_p2, lambda_q1, lambda_q2, eta;
    bool do_nmf, quiet, copy_data;
};

struct mf_model {            // mf/mf.h:70-79, 40 bytes; P,Q row-major [row][k], free()-able
    mf_int fun, m, n, k;
    mf_float b;
    mf_float *P, *Q;
};

// ---- on the accelerated path ------------------------------------------------------------------
MFB200_EXPORT mf_parameter mf_get_default_param();                                   // mf/mf.cpp:4538-4557
MFB200_EXPORT mf_model *mf_train(mf_problem const *prob, mf_parameter param);        // mf/mf.cpp:3362-3365
MFB200_EXPORT mf_model *mf_train_with_validation(mf_problem const *tr, mf_problem const *va,
                                                 mf_parameter param);                // mf/mf.cpp:3307-3332
MFB200_EXPORT mf_double mf_cross_validation(mf_problem const *prob, mf_int nr_folds,
                                            mf_parameter param);                     // mf/mf.cpp:4117-4129
MFB200_EXPORT mf_float mf_predict(mf_model const *model, mf_int u, mf_int v);        // mf/mf.cpp:4295-4314
MFB200_EXPORT mf_double calc_rmse(mf_problem *prob, mf_model *model);                // mf/mf.cpp:4316-4331
MFB200_EXPORT mf_double calc_mae(mf_problem *prob, mf_model *model);                 // mf/mf.cpp:4333-4347
MFB200_EXPORT mf_double calc_gkl(mf_problem *prob, mf_model *model);                 // mf/mf.cpp:4349-4364
MFB200_EXPORT mf_double calc_logloss(mf_problem *prob, mf_model *model);             // mf/mf.cpp:4366-4384
MFB200_EXPORT mf_double calc_accuracy(mf_problem *prob, mf_model *model);            // mf/mf.cpp:4386-4404
MFB200_EXPORT void mf_destroy_model(mf_model **model);                               // mf/mf.cpp:4280-4293
MFB200_EXPORT float *utility_train(float *train_data, int train_triplet_num, double p_l2, double q_l2,
                                   int k, int iters, double eta, int &lens);         // mf/mf.cpp:3483-3535
MFB200_EXPORT float *utility_predict(float *test_arr, int test_triplet_num, float *model_arr,
                                     int model_arr_len);                             // mf/mf.cpp:3537-3568
MFB200_EXPORT mf_problem read_triplet(float *tri, int triplet_num);                  // mf/mf.cpp:3367-3394
MFB200_EXPORT float *model_to_array(mf_model *model, int &lens);                     // mf/mf.cpp:3415-3441
MFB200_EXPORT mf_model *array_to_model(float *model_array, int lens);                // mf/mf.cpp:3444-3481

// ---- either side of the path: text formats (host) ---------------------------------------------
MFB200_EXPORT mf_problem read_problem(char const *path);                             // mf/mf.cpp:4143-4182
MFB200_EXPORT mf_int mf_save_model(mf_model const *model, char const *path);         // mf/mf.cpp:4184-4225
MFB200_EXPORT mf_model *mf_load_model(char const *path);                             // mf/mf.cpp:4227-4278
MFB200_EXPORT mf_int mf_my_train(char const *tr_path, char const *model_path);       // mf/mf.cpp:3397-3413

// ---- outside the path (SURVEY.md section 2 "OUT OF SCOPE"): exported, fail loudly --------------
MFB200_EXPORT float *cos_similarity(int item_id, float *q_arr, int q_arr_num);       // mf/mf.cpp:3591-3683
MFB200_EXPORT int *DINA(float *q_arr, int q_triplet_num, float *x_arr, int x_triplet_num,
                        int iterators);                                              // mf/mf.cpp:3685-4109
MFB200_EXPORT mf_model *mf_train_on_disk(char const *tr_path, mf_parameter param);   // mf/mf.cpp:4112-4115
MFB200_EXPORT mf_model *mf_train_with_validation_on_disk(char const *tr_path, char const *va_path,
                                                         mf_parameter param);        // mf/mf.cpp:3334-3360
MFB200_EXPORT mf_double mf_cross_validation_on_disk(char const *prob, mf_int nr_folds,
                                                    mf_parameter param);             // mf/mf.cpp:4131-4141
// the ranking measures of the one-class (BPR) losses, mf/mf.cpp:4406-4536, on the device (csrc/rank_metrics.cu); like the
// reference they sort prob->R in place
MFB200_EXPORT mf_double calc_mpr(mf_problem *prob, mf_model *model, bool transpose);
MFB200_EXPORT mf_double calc_auc(mf_problem *prob, mf_model *model, bool transpose);

}  // namespace mf

#endif  // MF_B200_HPP
