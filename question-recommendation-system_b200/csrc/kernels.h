/* csrc/kernels.h -- the thin C-ABI between the host engine (C++) and the sm_100a kernels.
 *
 * Everything here takes DEVICE pointers, plain sizes and a cudaStream_t passed as void*; each
 * function enqueues work on that stream and returns the cudaError_t of the launch as int (0 = ok).
 * Nothing synchronises unless its comment says so.  Host orchestration lives in engine.cpp.
 */
#ifndef MFB200_KERNELS_H
#define MFB200_KERNELS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfk_node {  /* == mf_node, mf/mf.h:36-41 */
    int u, v;
    float r;
} mfk_node;

/* Shape of the conflict-free band schedule.
 *
 * S side = the side with fewer rows ("stationary"): split into nC*nPass bands; CTA c keeps band
 * pass*nC+c in shared memory for a whole pass.  T side = the other side ("streaming"): split into
 * nTB = nC*S1 bands that rotate ring-wise over the CTAs; inside a band, group gamma of the CTA owns
 * T sub-band gamma.  A group is L lanes working on one rating.                                    */
typedef struct mfk_band_shape {
    int nC;         /* CTAs                                                                        */
    int nWarps;     /* warps per CTA                                                               */
    int L;          /* lanes per group: 8 (k_al <= 128) or 32                                      */
    int nG;         /* groups per CTA = nWarps * 32 / L                                            */
    int S1;         /* slack: T bands per CTA                                                      */
    int nTB;        /* T bands = nC * S1 = steps per pass                                          */
    int nPass;      /* passes per epoch (S bands per CTA)                                          */
    int segS;       /* rows per S band (<= rows_cap)                                               */
    int segT;       /* rows per T band                                                             */
    int segT2;      /* rows per T sub-band (group)                                                 */
    int rows_cap;   /* S rows that fit in shared memory                                            */
    int swap_sides; /* 1: S = users (m < n), 0: S = items                                          */
    int nStripes;   /* S side stripes (1 on one GPU; the number of ranks when the S side rotates)  */
    int stripeRows; /* S rows per stripe                                                           */
    int tLo, tRows; /* T rows owned by this rank: [tLo, tLo+tRows); others are dropped at load     */
    int tSeg;       /* T rows per rank (tRows of every rank but possibly the last)                 */
    int bitsA, bitsT, bitsD, bitsG, bitsSB, bitsB; /* key field widths: a_in, t, d, gamma, sb, b   */
    unsigned smem_bytes;
    int by_row;     /* order of the stream inside a (group, step) cell: 0 = phase, then T row (k_sgd_band_epoch);
                       1 = T row, so the ratings of a T row are adjacent (k_sgd_run_epoch, csrc/sgd_run.cu);
                       2 = no group field at all: the stream of an S band is ordered by (step, T row) and the offsets
                           are per (S band, step) -- the cells of k_sgd_cell_epoch (csrc/sgd_cell.cu)             */
    /*                 3 = as 1, but the T sub-bands belong to the WARPS (nG = nWarps units per CTA): the four groups
                           of a warp are served from one stream (k_sgd_warp_epoch, csrc/sgd_warp.cu)              */
    /*                 4 = the stream of a group is ordered by (S row, rotated T row): every S row belongs to one group for
                           the whole launch (k_sgd_item_epoch, csrc/sgd_item.cu); the step field holds S row / nG, w1 = S row */
    int chunk;      /* by_row == 2: entries a group takes off the CTA's cursor at a time (1..8)                  */
    int tlock;      /* by_row == 1, locks: 1 = T rows are taken one by one through lock words in global memory instead of
                       whole sub-bands through the ring's step hand-off (mfk_band_args.tlock)                       */
} mfk_band_shape;

/* rating stream word layouts */
#define MFK_W0_ABITS 20u /* w0 = t << 20 | a_in   (a_in: T row inside its band)                    */
#define MFK_W1_BBITS 13u /* w1 = ticket << 13 | b_local (S row inside its band)                    */
#define MFK_TICKET_MASK 0x7ffffu
/* cell stream (by_row == 2): w0 = HEAD | CONT | T row (relative to the rank's band), w1 = step << 13 | S row inside its
 * band.  HEAD: first entry of a run (the ratings of one T row inside a cell); CONT: the next entry belongs to the same run */
#define MFK_CELL_HEAD 0x80000000u
#define MFK_CELL_CONT 0x40000000u
#define MFK_CELL_ROW_MASK 0x3fffffffu

/* cross-validation (mf/mf.cpp:3208-3262): mask[nr_bins^2] on the device marks the hidden blocks of the reference's
 * nr_bins x nr_bins grid over the SHUFFLED ids (grid_problem, 793-858: seg_p = ceil(m / bins), seg_q = ceil(n / bins));
 * mask == NULL: nothing is hidden */
typedef struct mfk_hidden {
    const unsigned char *mask;
    int bins, seg_p, seg_q;
} mfk_hidden;

/* the loss codes of mf/mf.h:25-33 (mf_parameter.fun) that MFSolver implements */
enum { MFK_FUN_L2_MFR = 0, MFK_FUN_L1_MFR = 1, MFK_FUN_KL_MFR = 2, MFK_FUN_LR_MFC = 5, MFK_FUN_L2_MFC = 6,
       MFK_FUN_L1_MFC = 7, MFK_FUN_ROW_BPR = 10, MFK_FUN_COL_BPR = 11 /* one-class, exact mode only */ };

typedef struct mfk_band_args {
    float *S, *SG;            /* stationary side rows [nS][k_al] and AdaGrad accumulators [nS][2]  */
    float *T, *TG;            /* streaming side                                                    */
    const unsigned *w0, *w1;  /* rating stream, see above                                          */
    const float *rr;          /* ratings, already multiplied by 1/scale                            */
    const unsigned *goff;     /* [nC*nPass*nG + 1] first stream entry of every (S band, group) of
                                 the stripe this launch works on ([nC*nPass*nTB + 1], per (S band, step), when
                                 shape.by_row == 2)                                                */
    unsigned *flags;          /* [nC*nG] steps completed by every group, cumulative over launches
                                 (by_row == 2: [nC], per CTA)                                      */
    double *loss;             /* [1] += sum of e*e                                                 */
    int *error_flag;
    unsigned long long *stats; /* NULL, or 7 scheduling counters (tuning aid, see the kernel)         */
    mfk_band_shape shape;
    unsigned base;            /* cumulative step count at launch                                   */
    int nS;                   /* rows of the S side covered by this launch                         */
    int k_al;
    int dynamic;              /* 0: S rows by tickets (reproducible), 1: by locks (order depends on timing) */
    int late_lock;            /* with locks: take the lock after the T row arrived, release after the S row is stored */
    int full;                 /* 0: epoch 0, "slow only" (dims 0-7), mf/mf.cpp:2834,2910            */
    float lambda_s, lambda_t, eta;
    /* the general form of the update (any MFK_FUN_*, L1 regularisation, NMF); all zero = the L2_MFR fast path */
    int fun;
    float lambda1_s, lambda1_t;
    int do_nmf;
    double *err;              /* [1] += correctly classified ratings (the two hinge losses), may be NULL */
    unsigned long long wait_limit_ns; /* a hand-off wait longer than this (wall clock) makes the launch give up */
    unsigned *tlock;          /* run kernel, locks only: NULL = ring hand-off between CTAs; else one lock word per T row of
                                 this rank's band (all zero between launches): T rows are taken and returned one by one  */
} mfk_band_args;

int mfk_sm_count(int device);

/* sum r and r*r in double: out[0] += sum r, out[1] += sum r^2 (collect_info, mf/mf.cpp:462-484) */
int mfk_stats(const mfk_node *R, long long nnz, double *out2, void *stream);

/* band preprocessing (all on the device; replaces shuffle_problem 775-791, scale_problem 517-527
 * and grid_problem 793-858):
 *   keys1     remap ids, count omega, key1 = (stripe, S band) | group | step | phase | a_in (stream
 *             order), payload = b | r/scale; ratings of T rows outside [tLo, tLo+tRows) get key ~0
 *   sort64    stable LSD radix sort (cub)
 *   segstart  start of the (S band, group, step) segment of every entry -> rank inside the segment
 *   keys2     key2 = b | step | rank | group (ticket order of an S row), payload = stream position
 *   sort32
 *   tickets   ticket = position - first position of the row, scattered back to stream positions
 *   stream    w0, w1, rr and the per-(S band, group) offsets                                       */
int mfk_band_keys1(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, mfk_band_shape shape,
                   float inv_scale, int *omega_p, int *omega_q, unsigned long long *keys, unsigned long long *vals,
                   unsigned long long *kept_count, int *bad_index_flag, int m, int n, mfk_hidden hidden, void *stream);
/* sharded load (several GPUs): destination rank of every rating of a slice (+ omega on the slice, counts per
 * destination, out-of-range flag), and the grouping of the slice by destination                              */
int mfk_owner_of(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, int swap_sides, int t_seg,
                 int world, int *omega_p, int *omega_q, unsigned char *owner, unsigned long long *counts,
                 int *bad_index_flag, int m, int n, void *stream);
size_t mfk_group_tmp_bytes(long long n);
int mfk_group_by_owner(const unsigned char *owner_in, unsigned char *owner_out, const mfk_node *nodes_in,
                       mfk_node *nodes_out, long long n, int owner_bits, void *tmp, size_t tmp_bytes, void *stream);
size_t mfk_sort_tmp_bytes(long long n);
int mfk_sort_pairs32(unsigned long long *keys_in, unsigned long long *keys_out, unsigned *vals_in,
                     unsigned *vals_out, long long n, int end_bit, void *tmp, size_t tmp_bytes, void *stream);
int mfk_sort_pairs64(unsigned long long *keys_in, unsigned long long *keys_out, unsigned long long *vals_in,
                     unsigned long long *vals_out, long long n, int end_bit, void *tmp, size_t tmp_bytes,
                     void *stream);
int mfk_band_segstart(const unsigned long long *keys1_sorted, long long nnz, mfk_band_shape shape, unsigned *head,
                      unsigned *segstart, void *tmp, size_t tmp_bytes, void *stream);
int mfk_band_rank_bits(mfk_band_shape shape);
int mfk_band_keys2(const unsigned long long *keys1_sorted, const unsigned long long *vals1_sorted,
                   const unsigned *segstart, long long nnz, mfk_band_shape shape, unsigned long long *keys2,
                   unsigned *idx, void *stream);
int mfk_band_tickets(const unsigned long long *keys2_sorted, const unsigned *idx_sorted, long long nnz,
                     mfk_band_shape shape, unsigned *first, unsigned *ticket, void *stream);
int mfk_band_stream(const unsigned long long *keys1_sorted, const unsigned long long *vals1_sorted,
                    const unsigned *ticket, long long nnz, mfk_band_shape shape, unsigned *w0, unsigned *w1, float *rr,
                    unsigned *goff, void *stream);

/* init_model (mf/mf.cpp:952-1007) on the device.  rank[i] = number of rows j<i with omega[j]>0,
 * plus rank_base; draws are minstd_rand0 outputs number (rank*k + d + 1), jump-ahead computed.    */
int mfk_exclusive_rank(const int *omega, int rows, int *rank, int *total_out_dev, void *tmp,
                       size_t tmp_bytes, void *stream);
size_t mfk_rank_tmp_bytes(int rows);
int mfk_init_rows(float *M, float *G, const int *omega, const int *rank, int rank_base, int rows, int k,
                  int k_al, int zero_unseen, void *stream); /* zero_unseen: the BPR losses leave unseen rows 0 (997) */

/* the throughput kernel: one launch = one epoch of the band schedule (cooperative launch: CTAs wait
 * on one another, so all of them must be resident).  mfk_sgd_band_max_smem: usable dynamic shared memory. */
int mfk_sgd_band_epoch(const mfk_band_args *args, void *stream);
int mfk_sgd_band_max_smem(int device);
int mfk_sgd_band_max_warps(void); /* warps per CTA the band kernel was compiled for */
/* the run kernel (csrc/sgd_run.cu): same schedule and arguments, stream ordered by T row inside a cell (shape.by_row),
 * T rows kept in registers over a run and prefetched through shared memory; L2_MFR, k_al <= 128, 8 lanes per rating */
/* the item kernel (csrc/sgd_item.cu): shape.by_row == 4, T-row locks (args.tlock), one pass, no S band in shared memory */
unsigned mfk_sgd_item_smem_bytes(int k_al, int groups);
int mfk_sgd_item_epoch(const mfk_band_args *args, void *stream);
int mfk_sgd_item_max_warps(int lanes); /* lanes per group: 8 or 32 */
int mfk_sgd_run_max_warps(void); /* working warps per CTA the run kernel was compiled for */
int mfk_sgd_run_supported(int k_al, int L, int fun, float lambda1_s, float lambda1_t, int do_nmf);
unsigned mfk_sgd_run_slot_bytes(int k_al, int groups);
int mfk_sgd_run_epoch(const mfk_band_args *args, void *stream);
/* the cell kernel (csrc/sgd_cell.cu), for small launches: the CTA, not the group, owns a T band for a step and its groups
 * share the cell's ratings dynamically; shape.by_row == 2 (goff = per (S band, step) offsets, flags = one per CTA);
 * same support as the run kernel, locks only (args->dynamic != 0) */
/* the warp kernel (csrc/sgd_warp.cu): shape.by_row == 3, shape.nG == shape.nWarps (goff per (S band, warp), flags per
 * (CTA, warp)); same support as the run kernel, locks only; shared memory as the run kernel with 4 * nWarps groups */
int mfk_sgd_warp_epoch(const mfk_band_args *args, void *stream);
unsigned mfk_sgd_cell_extra_bytes(int k_al, int groups, int nTB);
int mfk_sgd_cell_epoch(const mfk_band_args *args, void *stream);

/* the exact kernel: one launch = one wavefront level of the reference's sequential order          */
int mfk_sgd_exact_level(const mfk_node *R, const unsigned *order, int count, float *P, float *Q, float *PG,
                        float *QG, int k_al, float lambda_p, float lambda_q, float eta, int slow_only,
                        float *e2_out, int fun, float lambda_p1, float lambda_q1, int do_nmf, float *err_out,
                        void *stream);
/* one-class BPR (BPRSolver, mf/mf.cpp:2131-2335), one wavefront level: visits [first, first+count) of order / neg   */
int mfk_bpr_exact_level(const mfk_node *R, const unsigned *order, const int *neg, int first, int count, float *P, float *Q,
                        float *PG, float *QG, int k_al, float lambda_p, float lambda_q, float eta, int slow_only,
                        float *loss_out, int col_oriented, float lambda_p1, float lambda_q1, int do_nmf, void *stream);
int mfk_sum_f32(const float *x, long long n, double *out1, void *stream); /* out[0] += sum (double) */

/* sum over rows with omega>0 of omega * <row,row> in SSE lane order (calc_reg2, 608-633)           */
int mfk_reg2(const float *M, const int *omega, int rows, int k_al, double *out1, void *stream);

/* sum over rows with omega>0 of omega * sum_d |row[d]| (calc_reg1, 583-606)                            */
int mfk_reg1(const float *M, const int *omega, int rows, int k_al, double *out1, void *stream);
/* the error measure of loss `which` (MFK_FUN_*): calc_mae/gkl/logloss/accuracy (4333-4404) with p_map == NULL,
 * calc_error (635-674) in training space with the permutations and 1/scale                              */
int mfk_err_general(int which, const mfk_node *R, long long nnz, const int *p_map, const int *q_map, const float *P,
                    const float *Q, int m, int n, int k, float b, float inv_scale, double *out1, int train_space,
                    mfk_hidden hidden, void *stream);
/* train_space: calc_error's form of the squared error; hidden.mask != NULL: only ratings of hidden blocks count (ids
 * after the permutations must then be training-space ids) and out1[1] += their number */

/* scale_model + shrink_model + shuffle_model (mf/mf.cpp:529-553,1057-1074,1027-1055):
 * out[id][0:k] = M[map[id]][0:k] * factor                                                          */
int mfk_finalize_rows(const float *M, const int *map, int rows, int k, int k_al, float factor, float *out,
                      void *stream);

/* mf_predict (4295-4314) per pair, exact sequential fp32 order; pairs as floats (3562-3565)        */
int mfk_predict_pairs(const float *P, const float *Q, int m, int n, int k, float b, const float *pairs,
                      long long npairs, float *out, void *stream);
/* calc_rmse's sum (4316-4331): out[0] += sum_double((float)(e*e))                                   */
int mfk_sq_err(const mfk_node *R, long long nnz, const float *P, const float *Q, int m, int n, int k,
               float b, double *out1, void *stream);
/* the validation column of the iteration table (mf/mf.cpp:2884-2904): out[0] += sum pow(r/scale - z, 2) in
 * double, z = mf_predict on the TRAINING-space model (stride k_al, ids through the permutations)            */
int mfk_va_err(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, const float *P, const float *Q,
               int m, int n, int k_al, float b, float inv_scale, double *out1, void *stream);

/* cosine similarity of the rows of an integer Q matrix (csrc/cos_sim.cu; mf::cos_similarity, mf/mf.cpp:3591-3683) for a
 * batch of items against all items, every row ordered by falling cosine; see the launcher for the buffers             */
size_t mfk_cos_tmp_bytes(int items, int a_count);
int mfk_cos_similarity(const int *Q, int items, int k, const int *a_list, int a_count, float *cos_raw, int *id_raw,
                       float *cos_sorted, int *id_sorted, int *tie_flags, void *tmp, size_t tmp_bytes, void *stream);

/* calc_mpr_auc (mf/mf.cpp:4406-4525) on the device (csrc/rank_metrics.cu): see the launchers for the buffers      */
size_t mfk_rank_sort_tmp_bytes(long long nnz);
int mfk_rank_prepare(const mfk_node *R, long long nnz, int transpose, int rows, unsigned long long *keys_tmp,
                     unsigned long long *keys_sorted, unsigned long long *kept_dev, long long *row_start, void *tmp,
                     size_t tmp_bytes, void *stream);
int mfk_rank_batch(const float *P, const float *Q, int m, int n, int k, float b, int transpose, int lo, int hi, int cols,
                   const unsigned long long *keys_sorted, const long long *row_start, long long first, long long last,
                   float *scores, float *pos_scores, float *pos_sorted, unsigned long long *u_mpr, unsigned long long *u_auc,
                   void *tmp, size_t tmp_bytes, void *stream);

/* Batched top-k of P.Q^T (csrc/topk.cu): bf16 tcgen05 GEMM passes + exact fp32 re-score; device pointers.
 * n <= 2048: every item is re-scored exactly (no GEMM).  Otherwise k <= 128 and topk <= 128 are required.
 * overflow_dev: int[1 + nusers] zeroed by the caller: [0] = number of users whose candidate list overflowed, then their
 * positions in `users`; mfk_topk_exact_user recomputes one such user exactly (every item scored, full sort) -- it is
 * also the path for k > 128 or topk > 128 with more than 2048 items (mfk_topk returns cudaErrorNotSupported).      */
size_t mfk_topk_exact_work_bytes(int n);
int mfk_topk_exact_user(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int pos, int topk,
                        int *idx_out, float *score_out, void *work, size_t work_bytes, void *stream);
int mfk_topk_max_candidates(void);
size_t mfk_topk_work_bytes(int n, int k, int batch_users, int sample_stride);
int mfk_topk(const float *P, const float *Q, int m, int n, int k, float b, const int *users, int nusers, int topk,
             int *idx_out, float *score_out, void *work, size_t work_bytes, int batch_users, int sample_stride,
             int sm_count, int *overflow_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif
