// csrc/engine.hpp -- host side of the training path (the replacement of fpsg / fpsg_core,
// mf/mf.cpp:2945-3042 and 2774-2943).  C++ host code; reaches the GPU only through kernels.h.
#ifndef MFB200_ENGINE_HPP
#define MFB200_ENGINE_HPP

#include <cstdint>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <string>
#include <utility>
#include <thread>
#include <vector>

#include "../../include/mfb200.h"
#include "kernels.h"

namespace mfb200 {

void set_error(const std::string &msg);  // thread-local last error + stderr
const char *last_error();

// Chooses the band schedule for a problem (DESIGN.md "choosing the shape").  nStripes > 1: the S side is
// cut into that many stripes that are trained one launch each (multi-GPU rotation); [t_lo, t_lo+t_rows)
// is the part of the T side this rank owns.  Returns false (with set_error) if the shape cannot be encoded.
bool plan_band(int m, int n, long long nnz, int k_al, int sm_count, int max_smem, int world, int rank,
               mfk_band_shape *out, int kernel = 0);  // kernel: 0 band, 1 run, 2 cell, 3 warp, 4 run or warp by launch size

// Rotation of the S stripes over the ranks.  With world > 1 the S side is cut into
// spr*world stripes (spr = stripes per rank); at sub-step sigma rank g trains stripe (spr*g + sigma) mod
// spr*world, then hands it to rank g-1, where it is needed spr sub-steps later (spr = 2: the transfer
// overlaps the next launch).
struct RotationStep {
    int compute;                  // half-stripe trained in this sub-step
    int send_stripe, send_to;     // after the sub-step
    int recv_stripe, recv_from;   // arrives during the next sub-step, used two sub-steps later
};
RotationStep rotation_step(int world, int rank, long long substep, int stripes_per_rank);
// sharded load: offsets of the all-to-all (counts[q * world + d] = ratings rank q holds for rank d) and the owner of a T row
long long exchange_plan(int world, int me, const unsigned long long *counts, long long *send_off, long long *recv_off);
int owner_of_row(int t_row, int t_seg, int world);

// pooled device memory (stream-ordered, legacy stream) and pipelined copies of pageable host arrays for the one-shot
// calls of mf_api.cpp; api_d2h returns when dst_host is complete
int api_pool_alloc(void **p, size_t bytes);
void api_pool_free(void *p);
int api_h2d(void *dst_dev, const void *src_host, size_t bytes);
int api_d2h(void *dst_host, const void *src_dev, size_t bytes);

std::vector<int> cv_block_order(int nr_blocks);  // do_cross_validation's shuffled block ids, mf/mf.cpp:3210-3214

// The two id permutations of a job (gen_random_map, mf/mf.cpp:1009-1017).  Several ranks inside ONE process (host threads,
// MFB200_GPUS) share one copy: the C library's rand() is process-wide state, so exactly one thread may walk it.
struct SharedMaps {
    std::once_flag once;
    std::vector<int> p, q;
};

class Session {
public:
    // rank/world/nccl_id: one process per GPU; nccl_id points to the 128-byte NCCL unique id of the job
    Session(int m, int n, const mfb200_param &prm, int rank = 0, int world = 1, const void *nccl_id = nullptr);
    ~Session();
    // validation set of mf_train_with_validation (host pointer, copied to the device by load())
    void set_validation(const mfb200_node *va, long long nnz) { va_host_ = va; va_nnz_ = va ? nnz : 0; }
    // cross-validation: grid blocks (ids of the reference's nr_bins x nr_bins grid) that are never trained on; call
    // before load().  cv_error: the loss's error measure over their ratings after training.
    void set_hidden_blocks(const int *blocks, int count);
    int cv_error(double *out);
    int load(const mfb200_node *R, long long nnz);
    int reset();
    int run_epochs(int epochs, float *ms_out, double *tr_rmse_out, bool print_table);
    int finish(float *P_out, float *Q_out, float *b_out);
    int heldout_rmse(const mfb200_node *R, long long nnz, double *out);
    void fill_report(mfb200_report *r) const;
    // ranks as host threads of one process: one of them generates the permutations, all use them
    void set_shared_maps(std::shared_ptr<SharedMaps> maps) { shared_maps_ = std::move(maps); }
    // a rank that computes the table's columns with the others (they are collective) but prints nothing
    void set_silent(bool silent) { silent_ = silent; }
    void release() { free_all(); }  // gives everything back now (the destructor then finds nothing to do)
    void *stream() const { return stream_; }
    int mode_used() const { return mode_; }

private:
    int init_device();
    static void *comm_for(int world, int rank, int device, const unsigned char *id128);  // cached ncclComm_t
    int init_model();
    void draw_block_generators();                        // BPR: Scheduler's per-block minstd_rand0, seeded from rand()
    int bpr_negative(int first_block, int second_block); // BPR: Scheduler::get_negative
    bool is_hidden(int blk) const;
    mfk_hidden hidden_arg() const;
    int upload_hidden_mask();
    int upload_maps();   // joins the helper thread that generates the permutations, copies them to the device
    int load_exact(const mfb200_node *R);
    int load_band(const mfb200_node *R);
    int epoch_exact(double *loss_out, double *err_out);
    int epochs_band(int epochs, double *loss_out, double *err_out);
    int finalize_to_device();
    int gather_model();   // world > 1: all-gather the T bands and the S stripes onto every rank
    void free_all();
    void print_header();
    void print_row(int iter, double tr_rmse, double va_rmse, double obj);
    int validation_error(double *va_rmse_out);
    int objective_terms(double *reg_out);

    int m_, n_, k_, k_al_;
    mfb200_param prm_;
    long long nnz_ = 0;
    int mode_ = 0, device_ = 0, sm_count_ = 0;
    int rank_ = 0, world_ = 1;
    unsigned char nccl_id_[128];
    void *comm_ = nullptr;          // ncclComm_t
    void *comm_stream_ = nullptr;   // all NCCL operations are issued on this stream
    std::vector<void *> kernel_done_, comm_done_;  // cudaEvent_t per sub-step of an epoch
    long long substeps_done_ = 0;
    bool reproducible_ = false;     // band mode: tickets instead of locks
    bool gathered_ = true;          // the full model is present on this rank
    size_t rowsP_alloc_ = 0, rowsQ_alloc_ = 0;
    void *stream_ = nullptr;
    void *ev0_ = nullptr, *ev1_ = nullptr;
    bool device_ready_ = false, loaded_ = false, header_printed_ = false;

    float avg_ = 0, std_dev_ = 0, scale_ = 1, lambda_p_ = 0, lambda_q_ = 0;
    float lambda_p1_ = 0, lambda_q1_ = 0;  // L1 coefficients after fpsg_core's rescaling
    int fun_ = 0;                          // MFK_FUN_* (mf_parameter.fun)
    bool bpr_ = false;                     // one of the two one-class BPR losses (exact mode only)
    std::vector<unsigned> block_gen_;      // BPR: the scheduler's per-block generators (negatives)
    int *d_neg_ = nullptr;                 // BPR: negative row of every visit of a portion
    int *h_neg_pinned_ = nullptr;
    bool regression_ = true;               // the three losses whose ratings are scaled by the standard deviation
    // layout of the small accumulator array: [0,1024) per-epoch loss sums, [1024,1040) scalars, then per-epoch error sums
    static constexpr int kAccErr = 1040, kAccSize = 1040 + 1024;
    std::vector<int> p_map_, q_map_;
    std::shared_ptr<SharedMaps> shared_maps_;
    bool silent_ = false;
    std::thread map_thread_;
    int epochs_done_ = 0;

    // device: training-space model
    float *dP_ = nullptr, *dQ_ = nullptr, *dPG_ = nullptr, *dQG_ = nullptr;
    int *d_omega_p_ = nullptr, *d_omega_q_ = nullptr, *d_pmap_ = nullptr, *d_qmap_ = nullptr;
    double *d_acc_ = nullptr;  // small accumulator array (pinned mirror h_acc_)
    double *h_acc_ = nullptr;
    int *d_err_ = nullptr;
    float *d_outP_ = nullptr, *d_outQ_ = nullptr;  // final-space model (stride k)

    // band (throughput) mode
    mfk_band_shape plan_{};
    long long nnz_kept_ = 0;        // ratings this rank trains on (== nnz_ on one GPU)
    unsigned *d_w0_ = nullptr, *d_w1_ = nullptr;
    float *d_rr_ = nullptr;
    unsigned *d_goff_ = nullptr, *d_flags_ = nullptr;
    unsigned *d_tlock_ = nullptr;  // run kernel with T-row locks: one word per T row of this rank's band, zero between launches
    unsigned step_base_ = 0;        // cumulative step count of the flags

    // exact mode
    mfk_node *d_R_ = nullptr;  // ratings in the reference's grid order
    unsigned *d_order_ = nullptr;
    float *d_e2_ = nullptr;
    std::vector<mfk_node> hR_;
    std::vector<long long> blk_first_;
    std::vector<int> visits_;
    typedef std::pair<float, int> Job;
    std::priority_queue<Job, std::vector<Job>, std::greater<Job>> heap_;
    std::default_random_engine sched_rng_;
    std::vector<int> lvl_u_, lvl_v_;
    std::vector<unsigned> order_host_;
    unsigned *h_order_pinned_ = nullptr;

    // cross-validation
    std::vector<int> hidden_;
    unsigned char *d_hidden_ = nullptr;
    mfk_node *d_cv_raw_ = nullptr;  // band mode: the raw ratings, kept for cv_error

    // validation set (training-space evaluation per epoch)
    const mfb200_node *va_host_ = nullptr;
    long long va_nnz_ = 0;
    mfk_node *d_va_ = nullptr;
    double last_va_rmse_ = 0;

    // report
    long long launches_ = 0;
    double prep_ms_ = 0, epochs_ms_ = 0, finish_ms_ = 0, last_tr_rmse_ = 0, create_ms_ = 0;
};

}  // namespace mfb200

#endif
