/* include/mfb200.h -- the C-ABI of the B200 matrix-factorisation engine.
 *
 * Plain C: pointers, sizes, PODs.  No C++ or torch types cross this boundary.  This is what a
 * foreign-function binding (PHP/Zend C glue, ctypes, cgo, JNI ...) links against.  Three groups:
 *
 *   1. php_*      the reference's own C-ABI (php_mf/mfWarp.h:6-10), same names and signatures, so
 *                 php_mf/php_mf.c (which declares them `extern` at php_mf.c:36-40) links unchanged.
 *   2. mfb200_*   one-shot calls with HOST buffers: the train / predict / metric path of mf/mf.h
 *                 without C++ by-value structs (mf/mf.h:89-91 passes mf_parameter by value and is
 *                 therefore not bindable from C).
 *   3. mfb200_session_*  the same training path split into stages, so a caller (bench.py) can keep
 *                 the ratings resident in HBM and time only the epoch loop.
 *
 * The C++ drop-in surface (namespace mf, Itanium-mangled, same struct layouts as mf/mf.h) is
 * declared in include/mf_b200.hpp and implemented by the same shared object (libmf.so).
 *
 * Error convention: functions returning int return 0 on success; non-zero on failure, with a
 * message retrievable through mfb200_last_error() (and printed on stderr, like the reference's
 * check_parameter does at mf/mf.cpp:3115-3184).  There is NO CPU fallback: without a usable
 * CUDA device every compute entry point fails.
 */
#ifndef MFB200_H
#define MFB200_H

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* exported even when the library is built -fvisibility=hidden */
#endif

/* ---- PODs (byte-identical to mf/mf.h:36-41 and the scalar part of mf/mf.h:51-66) ------------ */

typedef struct mfb200_node {  /* == mf::mf_node, mf/mf.h:36-41 */
    int u;                    /* row id   (user)     */
    int v;                    /* column id (question) */
    float r;                  /* rating              */
} mfb200_node;

enum {                         /* training modes (no counterpart in the reference)               */
    MFB200_MODE_AUTO = 0,      /* exact below MFB200_EXACT_MAX_NNZ ratings, else ring            */
    MFB200_MODE_EXACT = 1,     /* the reference's single-thread update order, bit-exact          */
    MFB200_MODE_RING = 2,      /* conflict-free parallel schedule (throughput); rows are handed out
                                  by locks, so the order of updates of a row depends on timing      */
    MFB200_MODE_RING_REPRO = 3 /* same schedule with rows handed out by tickets: bit-reproducible
                                  from run to run, ~15 % slower                                     */
};

typedef struct mfb200_param {  /* the knobs of mf_parameter that the training path reads             */
    int k;                     /* mf_parameter.k            (mf/mf.h:54)                          */
    int nr_bins;               /* mf_parameter.nr_bins      (mf/mf.h:56) -- exact mode only       */
    int nr_iters;              /* mf_parameter.nr_iters     (mf/mf.h:57)                          */
    float lambda_p2;           /* mf_parameter.lambda_p2    (mf/mf.h:59)                          */
    float lambda_q2;           /* mf_parameter.lambda_q2    (mf/mf.h:61)                          */
    float eta;                 /* mf_parameter.eta          (mf/mf.h:62)                          */
    int quiet;                 /* mf_parameter.quiet        (mf/mf.h:64): 0 prints the table      */
    int mode;                  /* MFB200_MODE_*                                                   */
    int device;                /* CUDA device ordinal (-1: current / env MFB200_DEVICE)           */
    /* appended (zero = the L2_MFR path above); see mf/mf.h:25-33 for the loss codes                  */
    int fun;                   /* mf_parameter.fun: 0 L2_MFR, 1 L1_MFR, 2 KL_MFR, 5 LR_MFC, 6 L2_MFC, 7 L1_MFC;
                                  10 ROW_BPR_MFOC, 11 COL_BPR_MFOC (one-class, exact mode on one device)  */
    float lambda_p1;           /* mf_parameter.lambda_p1    (mf/mf.h:58): L1 regularisation of P  */
    float lambda_q1;           /* mf_parameter.lambda_q1    (mf/mf.h:60)                          */
    int do_nmf;                /* mf_parameter.do_nmf       (mf/mf.h:63): project on x >= 0       */
} mfb200_param;

typedef struct mfb200_report {  /* filled by training calls; all times in milliseconds           */
    int mode_used;              /* MFB200_MODE_EXACT or MFB200_MODE_RING                          */
    int k_aligned;              /* ceil(k/8)*8, mf/mf.cpp:959                                     */
    int grid_ctas, cta_warps;   /* ring schedule shape (0 in exact mode)                          */
    int bands, subbands;        /* ring: column bands, sub-bands per band                         */
    long long launches;         /* number of SGD kernel launches                                  */
    double prep_ms;             /* H2D + preprocessing (stats, remap, scale, grid, init)          */
    double epochs_ms;           /* device time of the epoch loop (CUDA events)                    */
    double finish_ms;           /* un-scale, shrink, un-permute, D2H                              */
    double total_ms;            /* wall clock of the whole call (mfb200_train: create .. destroy) */
    double last_tr_rmse;        /* tr_rmse of the last epoch, as in the table mf/mf.cpp:2859-2867 */
    double create_ms;           /* device, stream, pool and pinned staging set-up (inside prep_ms
                                   for staged sessions: their first load() does it)               */
    double destroy_ms;          /* mfb200_train only: giving the device memory back to the pool   */
    int kernel;                 /* SGD kernel: 0 k_sgd_exact_level, 1 k_sgd_band_epoch, 2 k_sgd_run_epoch, 3 k_sgd_cell_epoch, 4 k_sgd_warp_epoch, 5 k_sgd_run_epoch with T-row locks */
    int gpus;                   /* devices the call trained on (1, or MFB200_GPUS for the one-shot calls)  */
} mfb200_report;

/* ---- library / device ------------------------------------------------------------------------ */

int mfb200_device_count(void);           /* number of usable CUDA devices (0: none -> train fails) */
const char *mfb200_last_error(void);     /* thread-local message of the last failure               */
const char *mfb200_version(void);
mfb200_param mfb200_default_param(void); /* mf_get_default_param, mf/mf.cpp:4538-4557             */

/* ---- group 1: the reference's C-ABI, php_mf/mfWarp.h:6-10 ------------------------------------ */

/* mf::mf_my_train (mf/mf.cpp:3397-3413): read "u v r" text, train 40 iters, save text model.     */
int php_mf_my_train(char *tr_path, char *model_path);
/* mf::utility_train (mf/mf.cpp:3483-3535): float triplets -> malloc'd float[5+m*k+n*k]; *lens.    */
float *php_utility_train(float *train_data, int train_triplet_num, double p_l2, double q_l2, int k,
                         int iters, double eta, int *lens);
/* mf::utility_predict (mf/mf.cpp:3537-3568): float pairs + model array -> malloc'd float[n].      */
float *php_utility_predict(float *test_arr, int test_triplet_num, float *model_arr, int model_arr_len);
/* mf::cos_similarity (mf/mf.cpp:3591-3683): float triplets (item, knowledge point, value) of the integer Q matrix ->
 * malloc'd float[items] = the item ids by falling cosine with item_id.  Cosines and sort on the device
 * (csrc/cos_sim.cu); equal cosines / zero rows in the reference's exchange-sort order.  Never NULL: on bad input a
 * message on stderr and a zeroed list (php_mf.c:1211 dereferences the result unchecked).                        */
float *php_cos_similarity(int item_id, float *q_arr, int q_arr_num);
/* mf::DINA (mf/mf.cpp:3685-4115): OUT OF SCOPE.  Exported so php_mf links; prints a message
 * and returns a zeroed malloc'd int[64] (php_mf.c:1281 reads 20 entries unchecked), never NULL.                 */
int *php_DINA(float *q_arr, int q_triplet_num, float *x_arr, int x_triplet_num, int iterators);

/* ---- group 2: one-shot calls, host buffers --------------------------------------------------- */

/* calc_mpr / calc_auc (mf/mf.cpp:4406-4536), the ranking measures of the one-class (BPR) losses: mean percentile rank
 * and area under the ROC curve of the positives (r > 0) of every row against all other columns; transpose: per item.
 * prob_m / prob_n: the problem's sizes (rows beyond the model score b).  (row, column) pairs must be distinct.       */
int mfb200_mpr_auc(const mfb200_node *R, long long nnz, int prob_m, int prob_n, const float *P, const float *Q, int m, int n,
                   int k, float b, int transpose, double *mpr_out, double *auc_out);

/* Cosine similarity of Q-matrix rows for a BATCH of items against all items (the reference answers one item per call,
 * mf/mf.cpp:3591-3683).  q_triplets: float (item, knowledge point, value) triplets as in
 * php_cos_similarity; item_ids[n_ids] (NULL / n_ids <= 0 with outputs given: every item).  Outputs, each [n_ids][items]
 * and optional: order_out = item ids by falling cosine (equal cosines: rising id; zero rows, whose cosine is 0/0,
 * last), cos_sorted_out = their cosines, cos_by_item_out = the cosines in item order; ties_out[n_ids] = 1 where a row
 * holds equal cosines or NaNs.  *items_out / *k_out = the matrix size (all outputs NULL: size query only).       */
int mfb200_cos_similarity(const float *q_triplets, int n_triplets, const int *item_ids, int n_ids, int *order_out,
                          float *cos_sorted_out, float *cos_by_item_out, int *ties_out, int *items_out, int *k_out);

/* mf_train (mf/mf.cpp:3362-3365, fpsg 2945-3042) for fun = P_L2_MFR.
 * R: nnz host nodes with 0 <= u < m, 0 <= v < n.  P_out[m*k], Q_out[n*k] (stride k, original ids;
 * rows never rated are NaN, mf/mf.cpp:996-999), *b_out = mean rating.  report may be NULL.        */
int mfb200_train(const mfb200_node *R, long long nnz, int m, int n, const mfb200_param *param,
                 float *P_out, float *Q_out, float *b_out, mfb200_report *report);

/* utility_predict's loop (mf/mf.cpp:3562-3565) over mf_predict (4295-4314): pairs are floats
 * truncated to int; out-of-range or NaN rows -> b; dot product in sequential fp32 order.          */
int mfb200_predict_pairs(const float *P, const float *Q, int m, int n, int k, float b,
                         const float *pairs, long long npairs, float *out);

/* calc_rmse (mf/mf.cpp:4316-4331): sqrt( sum_double( (float)(e*e) ) / nnz ).                       */
int mfb200_rmse(const mfb200_node *R, long long nnz, const float *P, const float *Q, int m, int n,
                int k, float b, double *rmse_out);

/* calc_mae / calc_gkl / calc_logloss / calc_accuracy (mf/mf.cpp:4333-4404): `which` is the loss code whose error
 * measure is wanted (1 mae, 2 gkl, 5 logloss, 6 or 7 accuracy; 0 rmse).                                          */
int mfb200_metric(int which, const mfb200_node *R, long long nnz, const float *P, const float *Q, int m, int n,
                  int k, float b, double *out);

/* mf_cross_validation (mf/mf.cpp:4117-4129): the blocks of the reference's nr_bins x nr_bins grid are dealt to
 * nr_folds folds in the reference's shuffled order; each fold trains with its blocks hidden and measures the loss's
 * error (rmse / mae / gkl / logloss / accuracy) on them.  fold_errors[nr_folds] may be NULL; *mean_out = their mean. */
int mfb200_cross_validation(const mfb200_node *R, long long nnz, int m, int n, const mfb200_param *param, int nr_folds,
                            double *fold_errors, double *mean_out);

/* Top-k per user over all n items (score = mf_predict, order score desc, id asc).
 * idx_out[nusers*topk] (-1 padded when n < topk), score_out[nusers*topk] or NULL.                  */
int mfb200_topk(const float *P, const float *Q, int m, int n, int k, float b, const int *users,
                int nusers, int topk, int *idx_out, float *score_out);
/* device time (ms, CUDA events) of the scoring inside the calling thread's last mfb200_topk, copies excluded */
double mfb200_topk_last_ms(void);
/* A user whose candidate list overflows in the tensor-core path (many items tied at the cut: an all-zero user row,
 * duplicate item rows) is recomputed exactly -- every item scored like mf_predict, full sort -- and so is every user when
 * the shape is outside that path (k > 128 or topk > 128 with more than 2048 items): slower, never inexact.         */

/* ---- group 2b: a model resident on the device --------------------------------------------------
 * mf::utility_predict re-parses and re-copies the whole model on every call (mf/mf.cpp:3559); the one-shot calls above
 * upload P and Q every time.  A handle keeps the factors in HBM across calls (one upload, any number of predict /
 * metric / top-k calls); it belongs to the device that was current (or MFB200_DEVICE) when it was made.          */
typedef struct mfb200_model mfb200_model;
mfb200_model *mfb200_model_upload(const float *P, const float *Q, int m, int n, int k, float b); /* NULL on failure */
void mfb200_model_free(mfb200_model *model);
int mfb200_model_predict_pairs(const mfb200_model *model, const float *pairs, long long npairs, float *out);
int mfb200_model_metric(const mfb200_model *model, int which, const mfb200_node *R, long long nnz, double *out);
int mfb200_model_topk(const mfb200_model *model, const int *users, int nusers, int topk, int *idx_out, float *score_out);
/* device time (ms, CUDA events) of the kernel inside the calling thread's last predict / metric call, copies excluded */
double mfb200_eval_last_ms(void);

/* Synthetic ratings of SURVEY.md 8d (counter based): writes count nodes starting at index first.  */
void mfb200_gen_ratings(unsigned long long seed, int m, int n, long long first, long long count,
                        mfb200_node *out);

/* ---- group 3: staged training (ratings resident in HBM between stages) ----------------------- */

typedef struct mfb200_session mfb200_session;

mfb200_session *mfb200_session_create(int m, int n, const mfb200_param *param);
/* Validation set of mf_train_with_validation (mf/mf.cpp:3307-3332): call BEFORE _load; the host array must stay
 * valid until _load returns.  With it, a non-quiet run prints the va_rmse column (mf/mf.cpp:2884-2904).        */
int mfb200_session_set_validation(mfb200_session *s, const mfb200_node *va_host, long long nnz);
/* H2D of the ratings + all preprocessing of fpsg (mf/mf.cpp:2994-3016) on the device.             */
int mfb200_session_load(mfb200_session *s, const mfb200_node *R_host, long long nnz);
/* Re-initialise the factors and schedule state without re-uploading (for repeated timing).        */
int mfb200_session_reset(mfb200_session *s);
/* Run `epochs` more epochs of fpsg_core's loop (mf/mf.cpp:2848-2914).  ms_out (may be NULL) gets
 * the device time of those epochs measured with CUDA events on the engine's stream;
 * tr_rmse_out (may be NULL) gets `epochs` values of the table's tr_rmse column.                   */
int mfb200_session_epochs(mfb200_session *s, int epochs, float *ms_out, double *tr_rmse_out);
/* scale_model + shrink_model + shuffle_model (mf/mf.cpp:3032-3034) and D2H.  P_out / Q_out may be NULL: the final
 * model then stays on the device (several ranks: a collective call all the same; only the ranks that pass buffers
 * pay for the download).                                                                            */
int mfb200_session_finish(mfb200_session *s, float *P_out, float *Q_out, float *b_out);
/* Held-out RMSE with the CURRENT factors, ratings given in original ids on the host.              */
int mfb200_session_rmse(mfb200_session *s, const mfb200_node *R_host, long long nnz, double *rmse_out);
int mfb200_session_report(mfb200_session *s, mfb200_report *report);
/* CUDA stream the engine launches on (cudaStream_t as void*), for external event timing.          */
void *mfb200_session_stream(mfb200_session *s);
void mfb200_session_destroy(mfb200_session *s);

/* ---- group 4: more than one GPU (one process per GPU, NCCL over NVLink) ------------------------ */
/* The T side (the factor matrix with more rows) is partitioned over the ranks; the S side is cut into
 * stripes that rotate ring-wise (ncclSend/ncclRecv on a second stream; one stripe per rank by default,
 * MFB200_STRIPES_PER_RANK=2 overlaps the transfer with the next launch).  Every rank passes the WHOLE rating set to mfb200_session_load and keeps its share.
 * mfb200_session_epochs / _finish / _rmse are collective: all ranks must call them in the same order;
 * every rank ends up with the full model.  NCCL is loaded with dlopen on first use.                  */
/* The NCCL communicator of a (world, rank, device) is created by the first session of the process and reused by
 * later ones (their id128 is then ignored): creating it costs seconds, a worker process trains many models.       */
int mfb200_dist_unique_id(unsigned char id128[128]);     /* rank 0: ncclGetUniqueId; ship it to all ranks */
mfb200_session *mfb200_dist_session_create(int m, int n, const mfb200_param *param, int rank, int world,
                                           const unsigned char id128[128]);
/* the rotation schedule (host logic, no GPU needed): out5 = {compute, send_stripe, send_to, recv_stripe,
 * recv_from} for sub-step `substep` (counted over the whole run) on `rank`.                          */
void mfb200_dist_rotation(int world, int rank, long long substep, int stripes_per_rank, int out5[5]);
/* the sharded load (host logic, no GPU needed): every rank uploads a slice of the ratings and sends each rating to the
 * rank that owns its T row (T = the factor matrix with more rows; t_seg = rows per rank = out16[14] of _plan_band on
 * rank 0).  counts[q * world + d] = ratings rank q holds for rank d; send_off[d] / recv_off[q] = where rank `rank`'s
 * block for d starts in its grouped slice / where rank q's block lands; returns the number of ratings received.   */
long long mfb200_dist_exchange_plan(int world, int rank, const unsigned long long *counts, long long *send_off,
                                    long long *recv_off);
int mfb200_dist_owner_of_row(int t_row, int t_seg, int world);
/* the band schedule a problem would get (host logic, no GPU needed): out16 = {nC, nWarps, L, nG, S1, nTB,
 * nPass, segS, segT, segT2, swap_sides, nStripes, stripeRows, tLo, tRows, smem_bytes}; 0 on success.  */
int mfb200_plan_band(int m, int n, long long nnz, int k, int world, int rank, int sm_count, int max_smem,
                     int out16[16]);
/* which SGD kernel that plan launches (the codes of mfb200_report.kernel); -1 on failure */
int mfb200_plan_kernel(int m, int n, long long nnz, int k, int world, int rank, int sm_count, int max_smem);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MFB200_H */
