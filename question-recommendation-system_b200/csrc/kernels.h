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

/* Shape of the conflict-free ring schedule (DESIGN.md "ring schedule").                        */
typedef struct mfk_ring_shape {
    int nC;        /* CTAs == row bands; CTA c owns rows [c*segA1, (c+1)*segA1)                  */
    int nW;        /* warps per CTA == sub-row-bands per row band                                */
    int S1, S2;    /* slack strides: column bands nB1 = nC*S1, sub-bands per band nB2 = nW*S2    */
    int nB1, nB2;
    int segA1, segA2, segB1, segB2; /* band widths in rows / columns                             */
    int bitsA;     /* ceil(log2(nA)): low key bits                                               */
    long long nSub; /* nC*nB1*nW*nB2 sub-blocks                                                  */
} mfk_ring_shape;

typedef struct mfk_ring_args {
    float *A, *B;             /* factors of the owned side [nA][k_al] and the banded side [nB][k_al] */
    float *AG, *BG;           /* AdaGrad accumulators [rows][2], mf/mf.cpp:2835                    */
    const int *ra, *rb;       /* sorted ratings, structure of arrays                              */
    const float *rr;
    const unsigned *sub_off;  /* [nSub+1] first rating of every sub-block                          */
    unsigned *progress;       /* [nC] steps completed per CTA, monotone over the whole run         */
    double *loss;             /* [1] += sum of e*e of this epoch                                   */
    int *error_flag;          /* set non-zero if a wait timed out                                  */
    mfk_ring_shape shape;
    int k_al;
    int epoch;                /* 0-based; epoch 0 is "slow only" (dims 0-7), mf/mf.cpp:2834,2910   */
    float lambda_a, lambda_b, eta;
} mfk_ring_args;

int mfk_sm_count(int device);

/* sum r and r*r in double: out[0] += sum r, out[1] += sum r^2 (collect_info, mf/mf.cpp:462-484) */
int mfk_stats(const mfk_node *R, long long nnz, double *out2, void *stream);

/* ring preprocessing: remap ids (shuffle_problem 775-791), count omega (grid_problem 810-817),
 * build the sort key (sub-block id << bitsA | a) and the payload (original index).               */
int mfk_ring_keys(const mfk_node *R, long long nnz, const int *p_map, const int *q_map, int swap_sides,
                  mfk_ring_shape shape, int *omega_p, int *omega_q, unsigned long long *keys,
                  unsigned *vals, void *stream);
/* stable LSD radix sort of (key,val) pairs on key bits [0,end_bit); tmp sized by mfk_sort_tmp_bytes */
size_t mfk_sort_tmp_bytes(long long n);
int mfk_sort_pairs(unsigned long long *keys_in, unsigned long long *keys_out, unsigned *vals_in,
                   unsigned *vals_out, long long n, int end_bit, void *tmp, size_t tmp_bytes, void *stream);
/* gather sorted SoA ratings (r * inv_scale, scale_problem 517-527) and the sub-block offsets      */
int mfk_ring_gather(const mfk_node *R, long long nnz, const unsigned long long *keys_sorted,
                    const unsigned *vals_sorted, const int *p_map, const int *q_map, int swap_sides,
                    mfk_ring_shape shape, float inv_scale, int *ra, int *rb, float *rr,
                    unsigned *sub_off, void *stream);

/* init_model (mf/mf.cpp:952-1007) on the device.  rank[i] = number of rows j<i with omega[j]>0,
 * plus rank_base; draws are minstd_rand0 outputs number (rank*k + d + 1), jump-ahead computed.    */
int mfk_exclusive_rank(const int *omega, int rows, int *rank, int *total_out_dev, void *tmp,
                       size_t tmp_bytes, void *stream);
size_t mfk_rank_tmp_bytes(int rows);
int mfk_init_rows(float *M, float *G, const int *omega, const int *rank, int rank_base, int rows, int k,
                  int k_al, void *stream);

/* the throughput kernel: one launch = one epoch of the ring schedule (cooperative launch)          */
int mfk_sgd_ring_epoch(const mfk_ring_args *args, void *stream);

/* the exact kernel: one launch = one wavefront level of the reference's sequential order          */
int mfk_sgd_exact_level(const mfk_node *R, const unsigned *order, int count, float *P, float *Q, float *PG,
                        float *QG, int k_al, float lambda_p, float lambda_q, float eta, int slow_only,
                        float *e2_out, void *stream);
int mfk_sum_f32(const float *x, long long n, double *out1, void *stream); /* out[0] += sum (double) */

/* sum over rows with omega>0 of omega * <row,row> in SSE lane order (calc_reg2, 608-633)           */
int mfk_reg2(const float *M, const int *omega, int rows, int k_al, double *out1, void *stream);

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
/* same on the training-space model (stride k_al, shuffled ids, scaled ratings): used per epoch     */

#ifdef __cplusplus
}
#endif
#endif
