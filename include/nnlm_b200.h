/* nnlm_b200.h — C ABI of the B200-native ANLS (alternating non-negative least squares) hot path.
 *
 * This is the drop-in boundary for the two `.Call` routines of the NNLM R package:
 *
 *   nnlm_nnmf  replaces  c_nnmf   (reference src/nnmf.cpp:4-220, bound at src/RcppExports.cpp:30-54,
 *                                   called from R/RcppExports.R:8-10 <- R/nnmf.R:177-182)
 *   nnlm_nnlm  replaces  c_nnlm   (reference src/nnlm.cpp:4-53,  bound at src/RcppExports.cpp:11-27,
 *                                   called from R/RcppExports.R:4-6  <- R/nnlm.R:119-120)
 *   nnlm_update replaces update / update_with_missing (reference src/update_with_missing.cpp:3-55, 58-139;
 *                                   prototypes src/nnlm.h:38-44) — one half-iteration, exposed for parity tests.
 *
 * Conventions (all arrays column-major, exactly the memory R hands to `.Call`):
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary.
 *   - inputs are borrowed read-only for the duration of the call; W/H are in/out.
 *   - missing entries of A are any non-finite double (NA_real_, NaN, +-Inf): src/update_with_missing.cpp:80-83.
 *   - masks are R logical matrices (int32, non-zero = fixed entry) or NULL for "no mask"
 *     (reference: empty umat, src/nnmf.cpp:75-80).
 *   - every function returns 0 on success, a negative NNLM_E_* code on failure, and writes a
 *     NUL-terminated message into `err` (if errlen > 0). The R shim turns that into Rf_error().
 *
 * The same signatures (prefix `oracle_` instead of `nnlm_`) are implemented by the CPU oracle in
 * oracle/nnlm_oracle.cpp so tests can diff the two call-for-call. The oracle is test infrastructure
 * only; this library never calls it and has no CPU fallback: without a CUDA device every compute entry
 * point returns NNLM_E_NO_DEVICE.
 */
#ifndef NNLM_B200_H
#define NNLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNLM_B200_ABI_VERSION 2

/* status codes */
#define NNLM_OK             0
#define NNLM_E_ARG         -1   /* bad argument (dimension, NULL pointer, method code...)      */
#define NNLM_E_NO_DEVICE   -2   /* no CUDA device / driver: there is NO CPU fallback            */
#define NNLM_E_CUDA        -3   /* a CUDA runtime/driver call or a kernel failed                */
#define NNLM_E_NCCL        -4   /* NCCL failure on the sharded path                             */
#define NNLM_E_INTERRUPT   -5   /* the interrupt callback asked to stop (Rcpp::checkUserInterrupt, src/nnmf.cpp:111) */
#define NNLM_E_NOMEM       -6

/* method codes: R/misc.R:28-35, src/nnmf.cpp:28-32 */
#define NNLM_SCD_MSE 1
#define NNLM_LEE_MSE 2
#define NNLM_SCD_MKL 3
#define NNLM_LEE_MKL 4

/* precision policy of the device copy of A:
 *   1 exact : A kept in f64, every product and sum in fp64 (the reference's arithmetic type); results agree with the
 *             reference to ~1e-13 on well-conditioned problems.
 *   2 fast  : square-loss methods on dense A: A kept as two fp16 planes, (A - mean)*s = hi + lo*2^-11 (22-24 significant
 *             bits, 4 B/element), cross-products on tcgen05 (3 fp16 MMAs per step, fp32 TMEM accumulators drained into fp64
 *             every 1024 contraction indices); measured error of the cross-product 1e-10..1e-9 relative to sum|F||A|, of
 *             the factors after one iteration ~1e-7 (tests/test_gpu_cross.py, tests/test_gpu_scale_parity.py).
 *             KL methods and the NA path keep A in f32 (relative rounding 6e-8) with every product and sum in fp64.
 *             Solver state (Gram, mu, h) is fp64 in every mode.
 *   0 auto  : nnlm_nnmf / sessions: exact when n*m < 4e6; otherwise fast, except that dense square-loss problems whose
 *             largest entry exceeds rms(A) * min(n, m) / 16 (count-like data with isolated huge entries) stay exact — the
 *             fp32 accumulators lose the small products added after such a spike. nnlm_nnlm and nnlm_update: always exact
 *             (their callers ask for rel_tol down to 1e-12, R/nnlm.R:72).
 * The R shim reads NNLM_B200_PRECISION = exact | fast | auto (default auto).                                          */
#define NNLM_PREC_AUTO  0
#define NNLM_PREC_EXACT 1
#define NNLM_PREC_FAST  2

/* Called between half-iterations on the calling thread; return non-zero to abort (-> NNLM_E_INTERRUPT). */
typedef int (*nnlm_interrupt_fn)(void* user);

/* Sink for the verbose == 2 iteration table of src/nnmf.cpp:100-104,155-156,194-198 (the R shim passes a wrapper of
 * Rprintf so the text lands on R's console); NULL = stdout. Called on the calling thread only. */
typedef void (*nnlm_print_fn)(void* user, const char* text);

/* Optional knobs that have no counterpart in the reference signature. Zero-initialise for defaults. */
typedef struct nnlm_options {
    int32_t precision;        /* NNLM_PREC_*                                                       */
    int32_t device;           /* CUDA device ordinal (the first one when n_gpus > 1), -1 = current */
    int32_t verbose_timing;   /* 1 = fill the timing fields of nnlm_stats                          */
    int32_t n_gpus;           /* nnlm_nnmf: 0 = environment NNLM_B200_GPUS (default 1); N > 1 = shard this ONE call over the
                                 devices device .. device+N-1: the calling thread drives rank 0, N-1 worker threads the
                                 others, NCCL inside the library (SURVEY.md §8b "Threading"). Ignored when comm is set. */
    void*   comm;             /* nnlm_comm handle for the one-process-per-GPU sharded sessions, or NULL */
    nnlm_print_fn print;      /* verbose output sink, NULL = stdout                                */
    void*   print_user;
    int32_t mkl_trace;        /* 0 = like the reference: the KL distance is evaluated at every error record
                                 (src/nnmf.cpp:137-139). 1 = square-loss methods on dense A evaluate it at the FINAL record
                                 only (earlier records hold NaN): their target error needs only the MSE, which comes from
                                 quantities already on the device (Engine::errors), so tracing costs no pass over A. */
    int32_t reserved1;
} nnlm_options;

typedef struct nnlm_stats {
    double upload_ms;         /* host->device copy + layout conversion of A, W, H, masks           */
    double loop_ms;           /* device time of the outer ANLS loop (CUDA events)                  */
    double download_ms;       /* device->host copy of W, H                                         */
    double cross_ms;          /* device time in the cross-product kernels                          */
    double solve_ms;          /* device time in the per-column solver kernels                      */
    double error_ms;          /* device time in error evaluation                                   */
    uint64_t launches;        /* kernels of this library launched during the call                  */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    int32_t  precision_used;  /* NNLM_PREC_EXACT or NNLM_PREC_FAST                                 */
    int32_t  reserved0;
    double gram_ms;           /* device time in the Gram / row-sum / factor re-layout kernels      */
    uint64_t cross_launches;  /* launches summed into cross_ms (verbose_timing only)               */
    uint64_t solve_launches;
    double comm_ms;           /* device time in the NCCL collectives (sharded path)                */
    uint64_t comm_bytes;      /* bytes this rank received through them                             */
    double host_setup_ms;     /* host wall clock: allocation + upload + ingest + factor/mask set-up  */
    double host_loop_ms;      /* host wall clock of the outer loop (incl. error evaluations)        */
    double host_finish_ms;    /* host wall clock: download of W, H                                  */
    double host_total_ms;     /* host wall clock from entry to just before return (set-up + loop + finish + teardown) */
    int32_t n_gpus_used;      /* devices this call ran on                                           */
    int32_t mse_from_identity;/* 1 = the traced MSE came from ||A||^2 - 2<H,WtA> + <WtW,HHt> (no pass over A) */
    double host_alloc_ms;     /* host wall clock inside device allocations (cudaMallocAsync + sync) during the call */
    double host_teardown_ms;  /* host wall clock releasing the call's device state                                 */
} nnlm_stats;

/* ---- c_nnmf (src/nnmf.cpp:4-220) -------------------------------------------------------------
 * A      n x m, may contain non-finite = missing.
 * W      n x K  in: initial W (explicit; the shim draws the default 0.01*U(0,1) init of src/nnmf.cpp:84-87
 *               with the host RNG, masked entries zeroed), out: result (reference returns W.t(), :212).
 * H      K x m  in/out likewise (src/nnmf.cpp:92-98).
 * Wm     n x K  int32 or NULL;  Hm  K x m int32 or NULL.
 * alpha, beta   [L2, angle, L1] penalties on W / H (src/nnmf.cpp:20-21).
 * trace  < 1 is treated as 1 (src/nnmf.cpp:53).
 * mse, mkl, target, avg_epoch: caller-allocated, capacity err_cap >= ceil(max_iter/trace)+1 (src/nnmf.cpp:54-55);
 *        *n_err receives the number of valid entries (i_e, src/nnmf.cpp:200-206).
 * n_iter receives the outer iteration count i (src/nnmf.cpp:218).
 * converged receives 0 iff the reference would warn "Target tolerance not reached. Try a larger max.iter."
 *        (rel_err > rel_tol, src/nnmf.cpp:208-209; the shim applies show_warning).
 */
int nnlm_nnmf(const double* A, int64_t n, int64_t m, int32_t K,
              double* W, double* H, const int32_t* Wm, const int32_t* Hm,
              const double* alpha, const double* beta,
              uint32_t max_iter, double rel_tol, int32_t n_threads, int32_t verbose,
              uint32_t inner_max_iter, double inner_rel_tol, int32_t method, uint32_t trace,
              double* mse, double* mkl, double* target, double* avg_epoch, uint32_t err_cap,
              uint32_t* n_err, uint32_t* n_iter, int32_t* converged,
              nnlm_interrupt_fn interrupt, void* interrupt_user,
              const nnlm_options* opt, nnlm_stats* stats,
              char* err, size_t errlen);

/* ---- c_nnlm (src/nnlm.cpp:4-53): solve y = x beta, beta >= 0, one update() call --------------
 * x n x p, y n x q (y may contain missing), coef p x q in: beta0 (explicit), out: coefficients.
 * mask p x q int32 or NULL. n_iteration receives the summed sweep count (src/nnlm.cpp:42-51).
 */
int nnlm_nnlm(const double* x, const double* y, int64_t n, int64_t p, int64_t q,
              double* coef, const int32_t* mask, const double* alpha,
              uint32_t max_iter, double rel_tol, int32_t n_threads, int32_t method,
              int64_t* n_iteration,
              const nnlm_options* opt, nnlm_stats* stats,
              char* err, size_t errlen);

/* ---- update / update_with_missing (src/update_with_missing.cpp:3-55, 58-139) -----------------
 * One half-iteration: solve A ~ Wt' H for H >= 0 column by column, warm-started from H.
 * H k x m in/out; Wt k x n; A n x m; mask k x m int32 or NULL; beta[3].
 * with_missing: 0 = update(), 1 = update_with_missing(), -1 = choose like the callers do (any non-finite in A).
 * total_iter receives the summed per-column sweep count (the function's return value in the reference).
 */
int nnlm_update(double* H, const double* Wt, const double* A, const int32_t* mask, const double* beta,
                int32_t k, int64_t n, int64_t m,
                uint32_t max_iter, double rel_tol, int32_t n_threads, int32_t method, int32_t with_missing,
                int64_t* total_iter,
                const nnlm_options* opt, nnlm_stats* stats,
                char* err, size_t errlen);

/* ---- diagnostic: the cross-product Q = Wt * A (k x m) alone — the contraction `Wt * A.col(j)` of
 * src/update_with_missing.cpp:39,45 for all columns — through the kernels a half-iteration uses (opt->precision selects
 * the fp64 CUDA-core or the tcgen05 path). Non-finite entries of A are read as zero (the masked product of :91). */
int nnlm_cross(const double* Wt, const double* A, int32_t k, int64_t n, int64_t m, double* Q,
               const nnlm_options* opt, nnlm_stats* stats, char* err, size_t errlen);

/* ---- diagnostic: the per-column corrections of update_with_missing (src/update_with_missing.cpp:88-91) as the fast path
 * forms them, by ONE tensor-core contraction of the 0/1 missing mask with the Khatri-Rao self-product of the factor:
 *   S[j*width + a(a+1)/2 + b] = sum_{i : A[i,j] missing} Wt[a,i] Wt[b,i]  (a >= b),   S[j*width + k(k+1)/2 + a] = sum_{i missing} Wt[a,i]
 * so that WtW_j = Wt Wt' - S_j. width (returned) = k(k+1)/2 + k rounded up to a multiple of 128. */
int nnlm_na_corrections(const double* Wt, const double* A, int32_t k, int64_t n, int64_t m, double* S, int64_t s_capacity,
                        int64_t* width, const nnlm_options* opt, char* err, size_t errlen);

/* ---- device-resident session: the benchmark's "inputs already in HBM" path -------------------
 * nnlm_session_create uploads A once (the copy + layout conversion c_nnmf implies per call);
 * nnlm_session_run performs `iters` ANLS iterations (W-half then H-half, src/nnmf.cpp:109-161 without the
 * error bookkeeping) on the resident factors and reports device time from CUDA events on the library's stream;
 * nnlm_session_get/set move W (n x K) and H (K x m) in the nnlm_nnmf layout.
 */
typedef struct nnlm_session nnlm_session;

int nnlm_session_create(nnlm_session** out, const double* A, int64_t n, int64_t m, int32_t K,
                        const int32_t* Wm, const int32_t* Hm,
                        const double* alpha, const double* beta,
                        uint32_t inner_max_iter, double inner_rel_tol, int32_t method,
                        const nnlm_options* opt, char* err, size_t errlen);
/* Synthetic workload of BASELINE.md §4 generated directly in HBM (the matrix never exists on the host):
 *   A = u(base+1)(n x K) * u(base+2)(K x m) + noise * u(base+3),  entry NaN iff u(base+4) < na_frac,
 *   u(seed, idx) = (splitmix64(seed*0x9E3779B97F4A7C15 + idx) >> 11) * 2^-53, idx = i + n*j (global indices; with
 *   opt->comm every rank generates exactly its column and row shard). nnlm_synth_matrix writes the same matrix to host memory. */
int nnlm_session_create_synthetic(nnlm_session** out, int64_t n, int64_t m, int32_t K, uint64_t seed_base, double noise,
                                  double na_frac, const double* alpha, const double* beta,
                                  uint32_t inner_max_iter, double inner_rel_tol, int32_t method,
                                  const nnlm_options* opt, char* err, size_t errlen);
/* Sharded session (one process per GPU; opt->comm = this rank's nnlm_comm). Rank g of R holds
 *   Acol = A[:, c0 .. c0+mc)  (n x mc, column-major)   and   Arow = A[r0 .. r0+nr, :]  (nr x m, column-major)
 * with the equal-chunk bounds chunk = ceil(total/R), start = min(total, g*chunk) (nnlm_b200.shard.shard_bounds).
 * Factors are whole (n x K, K x m) on every rank. nnlm_synth_block writes any block of the synthetic matrix to host memory. */
int nnlm_session_create_sharded(nnlm_session** out, const double* Acol, const double* Arow, int64_t n, int64_t m, int32_t K,
                                const int32_t* Wm, const int32_t* Hm, const double* alpha, const double* beta,
                                uint32_t inner_max_iter, double inner_rel_tol, int32_t method,
                                const nnlm_options* opt, char* err, size_t errlen);
int nnlm_synth_block(double* A, int64_t n_global, int64_t row0, int64_t nr, int64_t col0, int64_t mc, int32_t k,
                     uint64_t seed_base, double noise, double na_frac, char* err, size_t errlen);
int nnlm_synth_matrix(double* A, int64_t n, int64_t m, int32_t k, int64_t col0, uint64_t seed_base, double noise,
                      double na_frac, char* err, size_t errlen);
int nnlm_session_set_factors(nnlm_session* s, const double* W, const double* H, char* err, size_t errlen);
int nnlm_session_get_factors(nnlm_session* s, double* W, double* H, char* err, size_t errlen);
int nnlm_session_run(nnlm_session* s, uint32_t iters, double* device_ms, int64_t* total_sweeps,
                     char* err, size_t errlen);
/* mse / mkl / target (with penalties) of the resident factors: src/nnmf.cpp:121-149, 224-240 */
int nnlm_session_error(nnlm_session* s, double* mse, double* mkl, double* target, char* err, size_t errlen);
/* the MSE alone. On the dense square-loss path, right after nnlm_session_run, it is formed from quantities the last
 * H-half left on the device, ||A||^2 - 2<H,WtA> + <WtW,HHt> (fp64, no pass over A; *from_identity = 1); otherwise it
 * falls back to the fused pass of nnlm_session_error (*from_identity = 0). SURVEY.md §8f-2, src/nnmf.cpp:135-140. */
int nnlm_session_mse(nnlm_session* s, double* mse, int32_t* from_identity, char* err, size_t errlen);
int nnlm_session_stats(nnlm_session* s, nnlm_stats* stats);
int nnlm_session_reset_stats(nnlm_session* s);   /* zero the timing / launch counters */
void nnlm_session_destroy(nnlm_session* s);

/* ---- multi-GPU plumbing (one process per GPU) ----------------------------------
 * The id is created on rank 0, moved to the other ranks by the host program (torch.distributed / MPI / a file),
 * and every rank calls nnlm_comm_init on its own device. NCCL is dlopen()ed on first use. */
#define NNLM_COMM_ID_BYTES 128
typedef struct nnlm_comm nnlm_comm;
int  nnlm_comm_unique_id(unsigned char id[NNLM_COMM_ID_BYTES], char* err, size_t errlen);
int  nnlm_comm_init(nnlm_comm** out, const unsigned char id[NNLM_COMM_ID_BYTES], int32_t rank, int32_t nranks,
                    int32_t device, char* err, size_t errlen);
void nnlm_comm_destroy(nnlm_comm* c);

/* ---- misc ------------------------------------------------------------------------------------ */
int nnlm_abi_version(void);
/* sizeof(nnlm_options) for which == 0, sizeof(nnlm_stats) for which == 1 (bindings check their mirrors against it) */
size_t nnlm_sizeof(int which);
/* number of visible CUDA devices (0 if none / no driver); fills name of device 0 when name != NULL */
int nnlm_device_count(char* name, size_t namelen);
/* bit-exact NA mask (src/update_with_missing.cpp:80-83,91): builds on the device the bit-plane
 * bit(i + n*j) = !isfinite(A[i + n*j]) packed LSB-first into 32-bit words over the column-major linear index,
 * and copies it back (words = ceil(n*m/32)). Also returns per-column missing counts (int64[m]) when non-NULL. */
int nnlm_na_mask(const double* A, int64_t n, int64_t m, uint32_t* bits, int64_t* col_missing,
                 char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* NNLM_B200_H */
