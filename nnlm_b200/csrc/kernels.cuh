// kernels.cuh — host-side launchers of the sm_100a kernels. One half-iteration of ANLS is, in the reference's
// vocabulary (src/update_with_missing.cpp:3-55): solve  A ~ Wt' H  for H >= 0, column by column, where
//   H  is k x ncol (in/out),  Wt is k x len,  A is len x ncol with each column contiguous.
// Both halves of an iteration use the same launchers with the roles swapped (src/nnmf.cpp:117-119,131-133):
//   H-half: (H, W,  A   : len = n, ncol = m)      W-half: (W, H, A^T : len = m, ncol = n).
#pragma once
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace nnlm {

// ---- ingest.cu: upload-time layout conversion (replaces the per-iteration A.t() copy of src/nnmf.cpp:117,131) ----
struct IngestStats {          // accumulated over chunks, all in fp64 / exact integers
    double  kl_const_sum;     // sum over finite a of (a+TINY)*log(a+TINY) - a      (src/nnmf.cpp:66-73)
    int64_t n_missing;        // number of non-finite entries                        (src/nnmf.cpp:64)
    double  sum_sq;           // sum over finite a of a^2 (the ||A||^2 term of the Gram-identity MSE, Engine::errors)
};
constexpr int INGEST_PART_WIDTH = 3;   // doubles per partial record of launch_ingest: kl_const, n_missing, sum_sq
// src: chunk of `jc` columns (col-major, leading dimension len) already on the device.
// dst_cm (len x ncol, may alias src's parent buffer -> pass nullptr to skip) and dst_rm (ncol x len) receive columns [j0, j0+jc).
// part: device scratch of at least ingest_part_count(len, jc) * INGEST_PART_WIDTH doubles.
template <typename TOut>
void launch_ingest(const double* src, int64_t len, int64_t ncol, int64_t j0, int64_t jc, TOut* dst_cm, TOut* dst_rm,
                   double* part, cudaStream_t st);
int64_t ingest_part_count(int64_t len, int64_t jc);
// deterministic fixed-order reduction of `count` partial records of `width` doubles -> out[width]
void launch_reduce_partials(const double* part, int64_t count, int width, double* out, cudaStream_t st);
// bit-plane of !isfinite(A) over the column-major linear index + per-column counts (src/update_with_missing.cpp:80-83)
void launch_na_bits(const double* A, int64_t len, int64_t ncol, uint32_t* bits, int64_t* col_missing, cudaStream_t st);
// transpose small factor matrices between the R layout (rows x K, col-major) and the device layout (K x rows)
void launch_transpose_d(const double* in, int64_t rows, int64_t cols, double* out, cudaStream_t st);
void launch_mask_to_u8_t(const int32_t* in, int64_t rows, int64_t cols, uint8_t* out_t, cudaStream_t st);   // (rows x cols) -> (cols x rows) bytes
void launch_mask_to_u8(const int32_t* in, int64_t count, uint8_t* out, cudaStream_t st);

// ---- synth.cu: device-side synthetic workload ----
void launch_uniform(double* out, int64_t count, uint64_t seed, uint64_t offset, double scale, cudaStream_t st);
// A (n x m local columns [col0, col0+m) of the global matrix) = u(base+1)(n x k) * u(base+2)(k x m_global) + noise*u(base+3);
// NaN iff u(base+4) < na_frac. Synchronises the stream.
void launch_synth(double* A, int64_t n, int64_t m, int k, int64_t col0, uint64_t base, double noise, double na_frac,
                  cudaStream_t st);
// the block rows [row0, row0+nr) x columns [col0, col0+mc) of the global n_global x (any) matrix, written with leading dimension nr
void launch_synth_block(double* A, int64_t n_global, int64_t row0, int64_t nr, int64_t col0, int64_t mc, int k, uint64_t base,
                        double noise, double na_frac, cudaStream_t st);

// ---- gram.cu: K1/K1r/K6 of SURVEY.md §2.1 ----
// G = Y Y' (k x k) with the reference's regularisation (src/update_with_missing.cpp:19-24). Y is k x len.
// part must hold gram_splits(len) * k * k doubles.
int  gram_splits(int64_t len);
void launch_gram(const double* Y, int k, int64_t len, const double* pen /*[3] host, nullptr = raw unregularised Gram*/, double* part, double* G, cudaStream_t st);
// G = Graw + the reference's regularisation (src/update_with_missing.cpp:20-24); used after the raw Gram was all-reduced
void launch_gram_regularise(const double* Graw, int k, const double* pen, double* G, cudaStream_t st);
// sumW = rowSums(Y) (src/update_with_missing.cpp:27); part must hold gram_splits(len) * k doubles
void launch_rowsum(const double* Y, int k, int64_t len, double* part, double* out, cudaStream_t st);

// ---- cross_simt.cu: K2 in fp64 on CUDA cores (the "exact" path) ----
// Qp[s][a + k*j] = sum over the s-th slice of i of Y[a + k*i] * A[i + len*j].   TA = double or float.
int  cross_simt_splits(int k, int64_t len, int64_t ncol);
template <typename TA>
void launch_cross_simt(const double* Y, const TA* A, int k, int64_t len, int64_t ncol, int splits, double* Qp, cudaStream_t st);

// ---- cross_tc.cu: K2 on tcgen05 (fp16 hi/lo planes, fp32 TMEM accumulate drained into fp64, stream-K) ----
struct CrossPlan {
    int k = 0, np = 0, grid = 0, slots = 0;
    bool pairs = false;     // tiles of 256 columns worked by CTA pairs (grid counts pairs)
    int64_t len = 0, ncol = 0, ld_a = 0, ld_f = 0, tiles = 0, kblocks = 0, units = 0;
};
bool      cross_tc_supported(int k);
int       cross_tc_np(int k);                 // padded rank (rows of the factor planes)
int64_t   cross_tc_ld(int64_t len);           // row pitch (elements) of a plane whose rows hold `len` contraction indices
CrossPlan cross_tc_plan(int k, int64_t len, int64_t ncol, bool pairs = false);
// Qp[slot][ncol][k] (slots = plan.slots, zero-filled here) = partial cross-products; sum over slots = F * A.
// a_* : planes of the A copy of this half (row j = column j of the matrix, pitch plan.ld_a); f_* : planes of the factor.
// center[ncol] = the per-column mean that was subtracted from A before the split, fsum[k] = rowSums(F): the mean component
// center[j]*fsum[a] is added back in fp64 (slot 0). drain: k-blocks of 64 indices per fp32 TMEM accumulation (0 = the default
// of the tile width: 4 for k <= 64, 1 above).
void launch_cross_tc(const CrossPlan& plan, const __half* a_hi, const __half* a_lo, const __half* f_hi, const __half* f_lo,
                     const double* unscale, const double* center, const double* fsum, double* Qp, cudaStream_t st, int drain = 0);
// means over the finite entries of the columns (ncol values) and rows (len values, nullptr to skip) of A
void launch_means(const double* A, int64_t len, int64_t ncol, double* colmean, double* rowmean, cudaStream_t st);
// maxbits[0] = bit pattern of max |A| over the finite entries (a u64: non-negative doubles order like integers, so shards
// combine with an integer max); sA[0] = the power-of-two scale derived from it
void launch_absmax(const double* A, int64_t total, unsigned long long* maxbits, cudaStream_t st);
void launch_scale_from_max(const unsigned long long* maxbits, double* sA, cudaStream_t st);
// A (len x ncol col-major fp64) -> planes of A (pitch ld_a) and of A' (pitch ld_t; pass nullptr to skip)
void launch_split_matrix(const double* A, int64_t len, int64_t ncol, const double* sA, const double* colmean,
                         const double* rowmean, __half* a_hi, __half* a_lo, int64_t ld_a,
                         __half* t_hi, __half* t_lo, int64_t ld_t, cudaStream_t st);
// F (k x len col-major fp64) -> planes [np][ld] with per-row power-of-two scales; unscale[a] = 1/(sA*scale[a])
void launch_split_factor(const double* F, int k, int64_t len, int64_t ld, int np, const double* sA, unsigned long long* rowmax,
                         double* scales, double* unscale, __half* hi, __half* lo, cudaStream_t st);
// The same plus fsum = rowSums(F) in two launches: one pass for row sums and maxima (finished by the last CTA), one for the
// planes. part: 2 * splits * k doubles (splits = gram_splits(len)); ticket: one zero-initialised counter owned by the caller.
void launch_factor_prep(const double* F, int k, int64_t len, int64_t ld, int np, const double* sA, double* part, int splits,
                        unsigned int* ticket, double* fsum, double* scales, double* unscale, __half* hi, __half* lo, cudaStream_t st);
// rowmax[a] = bit pattern of max_i |F[a,i]| (zeroed here)
void launch_rowmax(const double* F, int k, int64_t len, unsigned long long* rowmax, cudaStream_t st);
// exact contraction of integer-valued fp16 planes (cross_tc.cu MODE 1); plan from cross_tc_plan(128, len, ncol)
void launch_cross_tc_exact(const CrossPlan& plan, const __half* a_plane, const __half* f_plane, const double* unscale, double* Qp,
                           cudaStream_t st);
void launch_cross_tc_exact2(const CrossPlan& plan, const __half* a_plane, const __half* f0, const __half* f1, const double* unscale,
                            double* Qp, cudaStream_t st);
// the same by pairs of CTAs (tcgen05 cta_group::2) on 256-column tiles: plan from cross_tc_plan(128, len, ncol, true)
void launch_mask_tc2(const CrossPlan& plan, const __half* a_plane, const __half* f0, const __half* f1, const double* unscale,
                     double* Qp, cudaStream_t st);

// ---- solve_ls.cu: K3/K4/K5 — warp-per-column sequential coordinate descent / Lee multiplicative, square loss ----
// X (k x ncol) in/out; G regularised Gram (k x k); Qp split-K partials of Wt*A (splits x k x ncol); mask k x ncol bytes or null;
// l1 = beta(2); sweeps: device counter incremented by the summed sweep count (total_raw_iter).
void launch_solve_ls(int method, double* X, const double* G, const double* Qp, int splits, const uint8_t* mask,
                     int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps,
                     cudaStream_t st);

// ---- solve_scd_tpc.cu: K3/K4 thread-per-column SCD (method 1, dense A, k <= 64): same contract as launch_solve_ls ----
bool scd_tpc_supported(int k);
size_t scd_tpc_scratch_doubles();
void launch_scd_tpc(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol,
                    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, double* scratch, cudaStream_t st);

// ---- solve_ls_missing.cu: K9 + K4/K5, the NA path of the square loss (src/update_with_missing.cpp:58-117) ----
// Y k x len (the fixed factor), A len x ncol (non-finite = missing), Gfull the raw (unregularised) Gram of Y,
// Qp the split-K partials of the masked cross-product (missing entries read as zero), pen[3] the penalties.
template <typename TA>
void launch_solve_ls_missing(int method, double* X, const double* Y, const TA* A, const double* Gfull, const double* Qp,
                             int splits, const uint8_t* mask, int k, int64_t len, int64_t ncol, const double* pen,
                             unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st);

// ---- na_gram.cu: K9 on the tensor cores (fast-precision NA path) ----
constexpr int NA_SLICES = 4;     // 11-bit fixed-point slices per column of the Khatri-Rao product
constexpr int NA_TILE = 128;     // Z columns per contraction (rows of one factor-plane tile)
int     na_pair_count(int k);    // k (k + 1) / 2 pairs (a >= b) + k linear columns
int64_t na_packed_width(int k);  // na_pair_count rounded up to a multiple of NA_TILE: row pitch of the packed result
// fp16 0/1 planes of the missing indicator of A (len x ncol, column-major): a_plane[j][i] and t_plane[i][j]; either may be null
template <typename TA>
void launch_mask_planes(const TA* A, int64_t len, int64_t ncol, __half* a_plane, int64_t ld_a, __half* t_plane, int64_t ld_t,
                        cudaStream_t st);
// S[j][pair(a,b)] = sum_{i missing in column j} Y[a,i] Y[b,i] (a >= b, pair = a (a + 1) / 2 + b), S[j][k(k+1)/2 + a] =
// sum_{i missing} Y[a,i]; exact up to a 2^-44 truncation of each product against its column bound. plan = cross_tc_plan(128, len, ncol).
void launch_na_gram_tc(const CrossPlan& plan, const double* Y, int k, const __half* mask_plane, unsigned long long* rowmax,
                       __half* zplanes, double* zunscale, double* Qp, double* S, cudaStream_t st);
// the NA-path solve from the packed corrections: G_j = Gfull - S_j (+ regularisation), q_j = sum of the Qp slots
// (- center[j] * masked row sums when center != nullptr: the cross-product was formed on centred planes with missing = 0)
void launch_solve_ls_missing_packed(int method, double* X, const double* Gfull, const double* S, int64_t pt, const double* Qp,
                                    int splits, const double* center, const uint8_t* mask, int k, int64_t ncol, const double* pen,
                                    unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st);

// ---- solve_kl.cu: K7/K8, KL loss, dense and NA (src/base_algorithms.cpp:71-151, update_with_missing.cpp:119-131) ----
// Yr is the ROW-major copy of the fixed factor: Yr[c*len + i] = Y[c + k*i]; sumY = rowSums(Y) (launch_rowsum).
int    solve_kl_grid(int64_t ncol);
size_t solve_kl_scratch_doubles(int64_t len, int64_t ncol);
template <typename TA>
void launch_solve_kl(int method, double* X, const double* Yr, const TA* A, const double* sumY, const uint8_t* mask, int k,
                     int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol, int with_missing,
                     double* wh_scratch, unsigned long long* sweeps, cudaStream_t st);

// ---- solve_kl_fast.cu: K7/K8 with the column state on chip (fast-precision path: dense A in fp32, no missing entries) ----
// Y32 is the fp32 row-major copy of the fixed factor ([k][len], launch_factor_rows_f32); see the file header for the design.
bool solve_kl_fast_supported(int k, int64_t len);
void launch_factor_rows_f32(const double* Y, int k, int64_t len, float* out, cudaStream_t st);
// WH0 (nullable): len x ncol fp32, the product Y' X of the factors the solve starts from (launch_product_tc)
void launch_solve_kl_fast(int method, double* X, const float* Y32, const float* A, const float* WH0, const double* sumY,
                          const uint8_t* mask, int k,
                          int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol,
                          unsigned long long* sweeps, cudaStream_t st);

// ---- error_eval.cu: a8/a9 ----
// Sums over finite entries of A of (A - W'H)^2 and of -(A+TINY)*log(W'H+TINY) + W'H  (src/nnmf.cpp:121-141).
// W is k x n, H is k x m, A is n x m. out[0] = sum sq, out[1] = sum kl. part: scratch error_part_count()*2 doubles.
int64_t error_part_count(int64_t n, int64_t m);
template <typename TA>
void launch_error(const TA* A, const double* W, const double* H, int k, int64_t n, int64_t m, double* part, double* out,
                  cudaStream_t st);
// Gram-identity MSE: sum (A - W'H)^2 = ||A||^2 - 2 <H, WtA> + <WtW, HHt>, from quantities the
// H-half left on the device. out[0] = <X, sum over slots of Qp> (X k x ncol, Qp [slot][ncol][k]); part: stats_part_count doubles.
void launch_dot_factor_cross(const double* X, const double* Qp, int splits, int k, int64_t ncol, double* part, double* out,
                             cudaStream_t st);
void launch_dot_small(const double* a, const double* b, int count, double* out, cudaStream_t st);   // out[0] = <a, b>
// out[0] = sum X^2, out[1] = sum X, out[2] = sum_i (sum_a X[a,i])^2 = accu(X*X.t())   (src/nnmf.cpp:224-240)
int64_t stats_part_count(int64_t cols);

// ---- error_tc.cu: a8 on the tensor cores (fast-precision path) ----
bool error_tc_supported(int k);
int  error_tc_kp(int k);                       // padded rank of the row planes: 64 or 128
int  error_tc_grid(int64_t n, int64_t m);      // CTAs = partial records
// X (k x cols, column-major) -> fp16 planes [cols][kp] under one power-of-two scale; rs[0] = 1 / scale; maxbits: one u64 of scratch
void launch_split_rows(const double* X, int k, int64_t cols, __half* hi, __half* lo, float* rs, unsigned long long* maxbits,
                       cudaStream_t st);
// out[0] = sum (A - W'H)^2, out[1] = sum (A+e) g((W'H - A) / (A+e)) over the finite entries (see the file header)
void launch_error_tc(const float* A, int64_t n, int64_t m, int k, const __half* w_hi, const __half* w_lo, const float* rsw,
                     const __half* h_hi, const __half* h_lo, const float* rsh, double* part, double* out, cudaStream_t st);
// out[i + n * j] = sum_c W[c, i] H[c, j] in fp32 from the same planes
void launch_product_tc(int64_t n, int64_t m, int k, const __half* w_hi, const __half* w_lo, const float* rsw, const __half* h_hi,
                       const __half* h_lo, const float* rsh, float* out, cudaStream_t st);
void launch_factor_stats(const double* X, int k, int64_t cols, double* part, double* out, cudaStream_t st);

}  // namespace nnlm
