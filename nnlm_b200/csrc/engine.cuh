// engine.cuh — device-resident state of one ANLS problem and the half-iteration driver.
//
// Mirrors the reference's conventions (src/nnmf.cpp:75-98, 109-161): W is held transposed (k x n) for the whole run,
// both halves go through the same routine with the roles swapped,
//   W-half: update(W, H, A.t(), Wm, alpha)   (src/nnmf.cpp:117 / :131)
//   H-half: update(H, W, A,     Hm, beta)    (src/nnmf.cpp:119 / :133)
// The reference materialises A.t() every iteration; here the transposed copy is made once at upload (ingest.cu) and both
// copies stay resident in HBM (2 x n*m*elt bytes), so each half streams its own copy with unit stride along the
// contraction index.
//
// Sharding (one process per GPU). Rank g of R owns the columns [c0, c0+mc) of A for the H-half (columns
// of H are independent given W) and the rows [r0, r0+nr) of A for the W-half (rows of W are independent given H); its
// "column copy" is A[:, cols_g] and its "row copy" is A[rows_g, :]' — together 2 x n*m/R elements, the same footprint
// per GPU as the single-GPU layout divided by R. Every cross-product entry is formed entirely on one GPU. Exchanges per
// half-iteration: one all-reduce of the k x k Gram of the slice each rank just solved, and one all-gather of the solved
// factor slices (k x n or k x m doubles in total) so the next half sees the whole fixed factor. With R = 1 all of this
// degenerates to the single-GPU path (no communicator).
#pragma once
#include <vector>

#include "comm.cuh"
#include "kernels.cuh"

namespace nnlm {

enum class Storage { F64, F32, F16X2 };   // element type of the resident copies of A

struct ErrorTerms {     // raw sums, all fp64
    double sum_sq;      // sum over finite entries of (A - W'H)^2
    double sum_kl;      // sum over finite entries of -(A+TINY) log(W'H+TINY) + W'H
    double w_stats[3];  // sum W^2, sum W, accu(W W')   (src/nnmf.cpp:224-240)
    double h_stats[3];
};

// Per-kernel device timing (nnlm_options.verbose_timing): CUDA events recorded on the library's stream around the
// cross-product and solver launches; accumulated at collect().
class KernelTimer {
public:
    enum Cat { CROSS = 0, SOLVE = 1, GRAM = 2, ERROR = 3, COMM = 4, NCAT = 5 };
    ~KernelTimer();
    void enable(bool on) { on_ = on; }
    bool enabled() const { return on_; }
    void begin(Cat c, cudaStream_t st);
    void end(cudaStream_t st);
    void collect();                       // stream must be synchronised
    double ms[NCAT] = {0, 0, 0, 0, 0};
    uint64_t count[NCAT] = {0, 0, 0, 0, 0};
    void reset() { for (int i = 0; i < NCAT; i++) { ms[i] = 0; count[i] = 0; } }
private:
    struct Span { cudaEvent_t a, b; Cat c; };
    std::vector<Span> spans_;
    std::vector<cudaEvent_t> pool_;
    cudaEvent_t get();
    bool on_ = false;
    Cat cur_ = CROSS;
    cudaEvent_t cur_a_ = nullptr;
};

class Engine {
public:
    KernelTimer timer;
    // n, m: GLOBAL dimensions of A. both_sides = false: only the H-half will run (nnlm_update / nnlm_nnlm), no row copy
    // of A is kept. comm: communicator of the column/row-sharded path, or nullptr for one GPU.
    Engine(int64_t n, int64_t m, int k, int method, int precision, int device, bool both_sides = true, Comm* comm = nullptr);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    // the shard of this rank (the whole matrix on one GPU)
    int64_t row0() const { return r0_; }
    int64_t rows_local() const { return nr_; }
    int64_t col0() const { return c0_; }
    int64_t cols_local() const { return mc_; }

    // single GPU: A is the whole n x m column-major host matrix (borrowed for the duration of the call only)
    void upload_A(const double* A);
    // device-resident shards (fp64, column-major): Acol = A[:, col0 .. col0+mc) (n x mc), Arow = A[row0 .. row0+nr, :]
    // (nr x m, may be nullptr when both_sides is false). On one GPU both may point to the same n x m matrix.
    void ingest_shards(const double* dAcol, const double* dArow);
    void set_factors(const double* W /*n x k*/, const double* H /*k x m*/);      // whole factors, on every rank
    void get_factors(double* W, double* H);
    void set_factors_t(const double* Wt /*k x n*/, const double* H /*k x m*/);
    void get_H(double* H);
    // -1: choose like c_nnmf/c_nnlm (NA path iff A has a non-finite entry); 0 / 1: force update() / update_with_missing()
    void set_missing_mode(int mode) { missing_mode_ = mode; }
    bool use_missing_path() const { return missing_mode_ < 0 ? n_missing_ > 0 : missing_mode_ != 0; }
    void set_masks(const int32_t* Wm /*n x k or null*/, const int32_t* Hm /*k x m or null*/);
    void set_penalties(const double* alpha, const double* beta);
    void set_inner(unsigned max_iter, double rel_tol) { inner_max_iter_ = max_iter; inner_rel_tol_ = rel_tol; }

    void half_w();      // solve for W given H (this rank's rows), then all-gather W
    void half_h();      // solve for H given W (this rank's columns), then all-gather H
    // global sums; synchronises the stream. want_kl = false: the KL sum is not needed (sum_kl = NaN) — on the dense square-loss
    // path sum_sq then comes from ||A||^2 - 2<H, WtA> + <WtW, HHt> (quantities the H-half left on the device, fp64) instead
    // of a pass over A; it falls back to the fused pass whenever those quantities are not current.
    void errors(ErrorTerms* out, bool want_kl = true);
    bool last_mse_from_identity() const { return last_identity_; }
    double sum_sq_A() const { return sum_sq_a_; }
    // diagnostic: Q = Wt * A (k x m, missing entries of A read as zero) through the cross-product path of the current storage
    void cross_only(double* Q_host);
    // diagnostic: the per-column corrections of the NA path through the tensor-core mask contraction (na_gram.cu):
    // S_host[j * width + pair(a,b)] = sum_{i missing in column j} W[i,a] W[i,b], then the k masked row sums; returns width
    int64_t na_corrections(double* S_host);
    uint64_t take_sweeps();                       // read and reset the global total_raw_iter (synchronises)
    void sync();

    int64_t n() const { return n_; }
    int64_t m() const { return m_; }
    int k() const { return k_; }
    int method() const { return method_; }
    int nranks() const { return comm_ ? comm_->nranks() : 1; }
    bool any_missing() const { return n_missing_ > 0; }
    int64_t n_missing() const { return n_missing_; }      // global
    double kl_const_sum() const { return kl_const_sum_; } // global
    int precision_used() const { return storage_ == Storage::F64 ? NNLM_PREC_EXACT : NNLM_PREC_FAST; }
    cudaStream_t stream() const { return st_; }
    int device() const { return device_; }
    uint64_t h2d_bytes = 0, d2h_bytes = 0, comm_bytes = 0;

private:
    struct Half {
        double* X; int64_t ncol;      // solved slice (k x ncol)
        const double* Y; int64_t len; // fixed factor, whole (k x len)
        const double* Yloc; int64_t len_loc;   // the slice of Y this rank solved in the previous half (Gram partial)
        const void* A;                // len x ncol copy of A for this half (F64 / F32 storage)
        const uint8_t* mask;
        const double* pen;
        bool w_side;
    };
    void run_half(const Half& h);
    template <typename TA> void run_half_t(const Half& h);
    void run_half_tc(const Half& h);
    // G_ (regularised) and Graw_ from the all-reduced slice Grams, on the side stream: the solver is the only consumer, so
    // the Gram kernels and their all-reduce run next to the factor split / row sums / cross-product of the main stream
    void fork_gram(const Half& h, bool raw_only);
    void join_gram();
    void solve_dense_ls(const Half& h, int splits);
    void gather(double* full, int64_t chunk_cols);
    void ensure_scratch();

    int64_t n_, m_;
    int k_, method_, device_;
    bool both_sides_;
    Comm* comm_;
    int64_t chunk_n_, chunk_m_, r0_, nr_, c0_, mc_;
    int missing_mode_ = -1;
    Storage storage_;
    cudaStream_t st_ = nullptr, st2_ = nullptr;
    cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
    unsigned inner_max_iter_ = 50;
    double inner_rel_tol_ = 1e-9;
    double alpha_[3] = {0, 0, 0}, beta_[3] = {0, 0, 0};

    DevBuf<double> A64_, At64_;     // column copy n x mc and row copy m x nr, column-major
    DevBuf<float> A32_, At32_;
    // fp16 hi/lo planes for the tensor-core cross-product (cross_tc.cu): column copy [mc][ld(n)], row copy [nr][ld(m)],
    // factor [np][ld]
    DevBuf<__half> a_hi_, a_lo_, t_hi_, t_lo_, f_hi_, f_lo_;
    // NA path on the tensor cores (na_gram.cu): 0/1 planes of the missing indicator in both orientations, the fixed-point
    // slices of one tile of the Khatri-Rao product, and the packed per-column corrections
    DevBuf<__half> mk_, mkt_, zplanes_;
    DevBuf<double> zunscale_, Qp2_, S_;
    CrossPlan plan_na_h_, plan_na_w_;
    // tensor-core error evaluation (error_tc.cu): row planes of W and of this rank's columns of H, their inverse scales
    DevBuf<__half> ew_hi_, ew_lo_, eh_hi_, eh_lo_;
    DevBuf<float> rsw_, rsh_;
    DevBuf<double> scale_a_, fscales_, unscale_, colmean_, rowmean_;
    DevBuf<unsigned long long> rowmax_;
    CrossPlan plan_h_, plan_w_;
    int precision_req_ = NNLM_PREC_AUTO;
    DevBuf<float> Yr32_;            // fp32 row-major copy of the fixed factor (KL fast path)
    DevBuf<float> wh32_;            // len x ncol product of the factors a KL half starts from (KL fast path)
    DevBuf<double> Wt_, H_;         // whole factors, k x (chunk_n * R) and k x (chunk_m * R)
    DevBuf<uint8_t> Wm_, Hm_;       // k x n, k x m or empty
    bool has_wm_ = false, has_hm_ = false;
    int64_t n_missing_ = 0;
    double kl_const_sum_ = 0.0;
    double sum_sq_a_ = 0.0;         // ||A||^2 over the finite entries, global
    bool q_valid_ = false;          // Qp_ / Graw_ / q_slots_ describe the H-half that produced the current H
    int q_slots_ = 0;
    bool last_identity_ = false;

    // scratch
    DevBuf<double> gram_part_, rowsum_part_, G_, Graw_, G2_, sumY_, Qp_, Yr_, wh_, red_part_, small_, tpc_scratch_;   // small_: 16 doubles of results
    DevBuf<unsigned long long> sweeps_;
    // read-back area of the few scalars a call returns (sums, counters). Plain host memory on purpose: cudaMallocHost /
    // cudaFreeHost of even 128 bytes took up to 600 ms per call next to large pinned regions of the host program
    // (scratch/e2e_probe.py, the e2e variance of round 1); every read-back is followed by a stream synchronisation anyway.
    struct HostSmall { double v[16]; double* p = v; } host_small_;
};

}  // namespace nnlm
