// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
//
// CPU restatement (C++17 + OpenMP, IEEE double, column-major, no BLAS) of the ANLS hot path of the NNLM R
// package, written from the reference's algorithm, for use as the parity checker in tests/, in
// __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference` arm of bench.py. Nothing under
// nnlm_b200/ may import, link or call this file.
//
// Why a restatement: the reference cannot be built in this image (needs R, Rcpp, RcppArmadillo, RcppProgress and a
// BLAS; none present, no network — SURVEY.md §0, §8c). The dense products the reference delegates to Armadillo/BLAS
// (`Wt*Wt.t()`, `Wt*A.col(j)`, `WtW*H.col(j)`, `W.t()*H`, dot, sum; versions unpinned in DESCRIPTION:15-22) are
// plain IEEE-double linear algebra; only their summation order is unpinned (~1e-16 relative).
//
// Pinning (tests/test_oracle_golden.py): golden NNLS vector tests/testthat/test-nnlm.R:29-43, exact recovery
// test-nnlm.R:6-26, rank-3 reconstruction for all four methods test-nnmf.R:5-24, mask exactness + one NA :67-85,
// imputation of 10% missing :88-93, warning condition :57-58. W/H *values* are not pinned by any reference test
// ("parity unpinned" for W/H values) — only W·H, NNLS optima and the properties above are.
//
// Each function cites the reference lines it follows. Paths are relative to /root/reference.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double TINY = 1e-16;  // src/nnlm.h:17  TINY_NUM

// When non-zero, oracle_nnmf re-materialises the transposed copy of A before every W-half, as the reference does
// (`A.t()` bound to a const mat&, src/nnmf.cpp:117,131). The arithmetic is identical either way.
int g_faithful_transpose = 0;

struct Dims { int k; int64_t n; int64_t m; };

// ---- src/base_algorithms.cpp:3-37 ------------------------------------------------------------
// Gauss–Seidel coordinate descent on 1/2 h'Vh - h'(WtA - beta3); mu = V h - WtA (+beta3) is maintained.
int scd_ls(double* h, const double* V, double* mu, const int32_t* mask, int k, unsigned max_iter, double rel_tol)
{
    double rel_err = 1 + rel_tol;
    unsigned t = 0;
    for (; t < max_iter && rel_err > rel_tol; t++) {
        rel_err = 0;
        for (int c = 0; c < k; c++) {
            if (mask && mask[c] > 0) continue;
            double cand = h[c] - mu[c] / V[c + (size_t)k * c];
            if (cand < 0) cand = 0;
            if (cand == h[c]) continue;
            const double d = cand - h[c];
            const double* Vc = V + (size_t)k * c;
            for (int r = 0; r < k; r++) mu[r] += d * Vc[r];
            const double e = 2 * std::fabs(h[c] - cand) / (cand + h[c] + TINY);
            if (e > rel_err) rel_err = e;
            h[c] = cand;
        }
    }
    return (int)t;
}

// ---- src/base_algorithms.cpp:40-68 -----------------------------------------------------------
// Lee–Seung multiplicative rule applied coordinate after coordinate (each step sees the updated h).
int lee_ls(double* h, const double* V, const double* WtAj, double beta3, const int32_t* mask, int k,
           unsigned max_iter, double rel_tol)
{
    double rel_err = rel_tol + 1;
    unsigned t = 0;
    for (; t < max_iter && rel_err > rel_tol; t++) {
        rel_err = 0;
        for (int c = 0; c < k; c++) {
            if (mask && mask[c] > 0) continue;
            const double* Vc = V + (size_t)k * c;
            double den = 0;
            for (int r = 0; r < k; r++) den += Vc[r] * h[r];
            den += beta3;
            const double ratio = WtAj[c] / (den + TINY);
            h[c] *= ratio;
            const double e = 2 * std::fabs(ratio - 1) / (ratio + 1);
            if (e > rel_err) rel_err = e;
        }
    }
    return (int)t;
}

// ---- src/base_algorithms.cpp:71-116 ----------------------------------------------------------
// KL loss, quadratic approximation per coordinate. Wt is k x len (column-major), Aj has len entries.
// NOTE the reference adds beta(0) to `a` BEFORE forming a*h (lines 100-101); follow the code, not the vignette.
int scd_kl(double* h, const double* Wt, const double* Aj, const double* sumW, const int32_t* mask,
           const double* beta, int k, int64_t len, unsigned max_iter, double rel_tol, double* Ajt /*scratch len*/)
{
    double sumH = 0;
    for (int c = 0; c < k; c++) sumH += h[c];
    for (int64_t i = 0; i < len; i++) {                       // Ajt = Wt' h  (line 82)
        const double* w = Wt + (size_t)k * i;
        double s = 0;
        for (int c = 0; c < k; c++) s += w[c] * h[c];
        Ajt[i] = s;
    }
    double rel_err = 1 + rel_tol;
    unsigned t = 0;
    for (; t < max_iter && rel_err > rel_tol; t++) {
        rel_err = 0;
        for (int c = 0; c < k; c++) {
            if (mask && mask[c] > 0) continue;
            double a = 0, b = 0;
            for (int64_t i = 0; i < len; i++) {
                const double mu = Wt[c + (size_t)k * i] / (Ajt[i] + TINY);   // line 97
                a += Aj[i] * (mu * mu);                                       // dot(Aj, square(mu))
                b += Aj[i] * mu;
            }
            b -= sumW[c];
            a += beta[0];
            b += a * h[c] - beta[2] - beta[1] * (sumH - h[c]);
            double cand = b / (a + TINY);
            if (cand < 0) cand = 0;
            if (cand != h[c]) {
                const double d = cand - h[c];
                for (int64_t i = 0; i < len; i++) Ajt[i] += d * Wt[c + (size_t)k * i];
                const double e = 2 * std::fabs(h[c] - cand) / (cand + h[c] + TINY);
                if (e > rel_err) rel_err = e;
                sumH += cand - h[c];
                h[c] = cand;
            }
        }
    }
    return (int)t;
}

// ---- src/base_algorithms.cpp:119-151 ---------------------------------------------------------
int lee_kl(double* h, const double* Wt, const double* Aj, const double* sumW, const int32_t* mask,
           const double* beta, int k, int64_t len, unsigned max_iter, double rel_tol, double* wh /*scratch len*/)
{
    double sumH = 0;
    for (int c = 0; c < k; c++) sumH += h[c];
    for (int64_t i = 0; i < len; i++) {                       // wh = Wt' h (line 133)
        const double* w = Wt + (size_t)k * i;
        double s = 0;
        for (int c = 0; c < k; c++) s += w[c] * h[c];
        wh[i] = s;
    }
    double rel_err = rel_tol + 1;
    unsigned t = 0;
    for (; t < max_iter && rel_err > rel_tol; t++) {
        rel_err = 0;
        for (int c = 0; c < k; c++) {
            if (mask && mask[c] > 0) continue;
            double num = 0;
            for (int64_t i = 0; i < len; i++) num += Wt[c + (size_t)k * i] * (Aj[i] / (wh[i] + TINY));  // line 141
            double ratio = num / (sumW[c] + beta[0] * h[c] + beta[1] * (sumH - h[c]) + beta[2]);
            const double step = (ratio - 1) * h[c];
            for (int64_t i = 0; i < len; i++) wh[i] += step * Wt[c + (size_t)k * i];
            sumH += step;
            h[c] *= ratio;
            const double e = 2 * std::fabs(ratio - 1) / (ratio + 1);
            if (e > rel_err) rel_err = e;
        }
    }
    return (int)t;
}

// Gram G = Wt Wt' over the listed columns (all columns when idx == nullptr), then the reference's regularisation
// src/update_with_missing.cpp:19-24 (and :90,95,98-103).
void gram_regularised(double* G, const double* Wt, int k, int64_t n, const int64_t* idx, int64_t nidx, const double* beta)
{
    std::fill(G, G + (size_t)k * k, 0.0);
    const int64_t cnt = idx ? nidx : n;
    for (int64_t t = 0; t < cnt; t++) {
        const double* w = Wt + (size_t)k * (idx ? idx[t] : t);
        for (int c = 0; c < k; c++) {
            const double wc = w[c];
            double* Gc = G + (size_t)k * c;
            for (int r = 0; r < k; r++) Gc[r] += w[r] * wc;
        }
    }
    if (beta[0] != beta[1])
        for (int c = 0; c < k; c++) G[c + (size_t)k * c] += beta[0] - beta[1];
    if (beta[1] != 0)
        for (size_t e = 0; e < (size_t)k * k; e++) G[e] += beta[1];
    for (int c = 0; c < k; c++) G[c + (size_t)k * c] += TINY;
}

inline bool all_masked(const int32_t* mcol, int k)
{
    for (int c = 0; c < k; c++) if (!(mcol[c] > 0)) return false;
    return true;
}

int resolve_threads(int n_threads)
{
#ifdef _OPENMP
    if (n_threads <= 0) return omp_get_max_threads();   // src/update_with_missing.cpp:13 + R/nnmf.R:168
    return n_threads;
#else
    (void)n_threads; return 1;
#endif
}

// ---- src/update_with_missing.cpp:3-55 --------------------------------------------------------
// A is addressed as A[i*sa_i + j*sa_j] so the same code serves A (sa_i=1, sa_j=n) and the W-half's A.t() without
// a copy; the "faithful" benchmark variant passes an explicit transposed copy instead.
int64_t update_dense(double* H, const double* Wt, const double* A, int64_t sa_i, int64_t sa_j, const int32_t* mask,
                     const double* beta, Dims d, unsigned max_iter, double rel_tol, int n_threads, int method)
{
    const int k = d.k; const int64_t n = d.n, m = d.m;
    std::vector<double> G, sumW;
    if (method == 1 || method == 2) {
        G.resize((size_t)k * k);
        gram_regularised(G.data(), Wt, k, n, nullptr, 0, beta);
    } else {
        sumW.assign(k, 0.0);                                                   // sum(Wt, 1), line 27
        for (int64_t i = 0; i < n; i++) for (int c = 0; c < k; c++) sumW[c] += Wt[c + (size_t)k * i];
    }
    int64_t total = 0;
    const int nt = resolve_threads(n_threads);
    #pragma omp parallel num_threads(nt)
    {
        std::vector<double> mu(k), wta(k), scratch, acol;
        if (method >= 3) { scratch.resize(n); if (sa_i != 1) acol.resize(n); }
        #pragma omp for schedule(dynamic) reduction(+:total)
        for (int64_t j = 0; j < m; j++) {
            const int32_t* mcol = mask ? mask + (size_t)k * j : nullptr;
            if (mcol && all_masked(mcol, k)) continue;                          // lines 33-34
            double* h = H + (size_t)k * j;
            const double* Aj = A + j * sa_j;
            int iter = 0;
            if (method == 1 || method == 2) {
                std::fill(wta.begin(), wta.end(), 0.0);                         // Wt * A.col(j), lines 39,45
                for (int64_t i = 0; i < n; i++) {
                    const double a = Aj[i * sa_i];
                    const double* w = Wt + (size_t)k * i;
                    for (int c = 0; c < k; c++) wta[c] += w[c] * a;
                }
                if (method == 1) {
                    for (int r = 0; r < k; r++) {                               // mu = WtW*h - WtA (+beta2), 39-41
                        double s = 0;
                        for (int c = 0; c < k; c++) s += G[r + (size_t)k * c] * h[c];
                        mu[r] = s - wta[r];
                    }
                    if (beta[2] != 0) for (int r = 0; r < k; r++) mu[r] += beta[2];
                    iter = scd_ls(h, G.data(), mu.data(), mcol, k, max_iter, rel_tol);
                } else {
                    iter = lee_ls(h, G.data(), wta.data(), beta[2], mcol, k, max_iter, rel_tol);
                }
            } else {
                const double* col = Aj;
                if (sa_i != 1) { for (int64_t i = 0; i < n; i++) acol[i] = Aj[i * sa_i]; col = acol.data(); }
                if (method == 3) iter = scd_kl(h, Wt, col, sumW.data(), mcol, beta, k, n, max_iter, rel_tol, scratch.data());
                else             iter = lee_kl(h, Wt, col, sumW.data(), mcol, beta, k, n, max_iter, rel_tol, scratch.data());
            }
            total += iter;                                                      // lines 51-52
        }
    }
    return total;
}

// ---- src/update_with_missing.cpp:58-139 ------------------------------------------------------
int64_t update_missing(double* H, const double* Wt, const double* A, int64_t sa_i, int64_t sa_j, const int32_t* mask,
                       const double* beta, Dims d, unsigned max_iter, double rel_tol, int n_threads, int method)
{
    const int k = d.k; const int64_t n = d.n, m = d.m;
    int64_t total = 0;
    const int nt = resolve_threads(n_threads);
    #pragma omp parallel num_threads(nt)
    {
        std::vector<double> G((size_t)k * k), mu(k), wta(k), sumW(k), sub_w, sub_a, scratch;
        std::vector<int64_t> nm;
        #pragma omp for schedule(dynamic) reduction(+:total)
        for (int64_t j = 0; j < m; j++) {
            const int32_t* mcol = mask ? mask + (size_t)k * j : nullptr;
            if (mcol && all_masked(mcol, k)) continue;                          // lines 77-78
            double* h = H + (size_t)k * j;
            const double* Aj = A + j * sa_j;
            nm.clear();
            for (int64_t i = 0; i < n; i++) if (std::isfinite(Aj[i * sa_i])) nm.push_back(i);   // find_finite, 80-83
            const bool any_missing = (int64_t)nm.size() != n;
            const int64_t cnt = (int64_t)nm.size();
            int iter = 0;
            if (method == 1 || method == 2) {
                // per-column Gram (recomputed even for complete columns, lines 95-96) and cross-product 90-91
                gram_regularised(G.data(), Wt, k, n, any_missing ? nm.data() : nullptr, cnt, beta);
                std::fill(wta.begin(), wta.end(), 0.0);
                for (int64_t t = 0; t < cnt; t++) {
                    const int64_t i = nm[t];
                    const double a = Aj[i * sa_i];                              // A.elem(j*n + non_missing)
                    const double* w = Wt + (size_t)k * i;
                    for (int c = 0; c < k; c++) wta[c] += w[c] * a;
                }
                if (method == 1) {
                    for (int r = 0; r < k; r++) {                               // mu = WtW*h - mu (+beta2), 109-111
                        double s = 0;
                        for (int c = 0; c < k; c++) s += G[r + (size_t)k * c] * h[c];
                        mu[r] = s - wta[r];
                    }
                    if (beta[2] != 0) for (int r = 0; r < k; r++) mu[r] += beta[2];
                    iter = scd_ls(h, G.data(), mu.data(), mcol, k, max_iter, rel_tol);
                } else {
                    iter = lee_ls(h, G.data(), wta.data(), beta[2], mcol, k, max_iter, rel_tol);
                }
            } else {
                // KL methods receive the gathered sub-matrices Wt.cols(nm), A.elem(...), sum(Wt.cols(nm),1): 119-131
                sub_w.resize((size_t)k * cnt); sub_a.resize(cnt); scratch.resize(cnt);
                std::fill(sumW.begin(), sumW.end(), 0.0);
                for (int64_t t = 0; t < cnt; t++) {
                    const int64_t i = nm[t];
                    sub_a[t] = Aj[i * sa_i];
                    for (int c = 0; c < k; c++) { const double w = Wt[c + (size_t)k * i]; sub_w[c + (size_t)k * t] = w; sumW[c] += w; }
                }
                if (method == 3) iter = scd_kl(h, sub_w.data(), sub_a.data(), sumW.data(), mcol, beta, k, cnt, max_iter, rel_tol, scratch.data());
                else             iter = lee_kl(h, sub_w.data(), sub_a.data(), sumW.data(), mcol, beta, k, cnt, max_iter, rel_tol, scratch.data());
            }
            total += iter;                                                      // lines 135-136
        }
    }
    return total;
}

// ---- src/nnmf.cpp:224-240 --------------------------------------------------------------------
// W is k x n here, so accu(W*W.t()) = sum of all entries of the k x k Gram = sum_i (sum_a W[a,i])^2.
double penalty(const double* W, const double* H, Dims d, double N_non_missing, const double* alpha, const double* beta)
{
    auto sq = [](const double* X, size_t len) { double s = 0; for (size_t e = 0; e < len; e++) s += X[e] * X[e]; return s; };
    auto sm = [](const double* X, size_t len) { double s = 0; for (size_t e = 0; e < len; e++) s += X[e]; return s; };
    auto gs = [](const double* X, int k, int64_t cols) {
        // accu(X * X.t()) for X k x cols
        double s = 0;
        for (int a = 0; a < k; a++) for (int b = 0; b < k; b++) {
            double g = 0;
            for (int64_t i = 0; i < cols; i++) g += X[a + (size_t)k * i] * X[b + (size_t)k * i];
            s += g;
        }
        return s;
    };
    double p = 0;
    const size_t lw = (size_t)d.k * d.n, lh = (size_t)d.k * d.m;
    if (alpha[0] != alpha[1]) p += 0.5 * (alpha[0] - alpha[1]) * sq(W, lw) / N_non_missing;
    if (beta[0] != beta[1])   p += 0.5 * (beta[0] - beta[1]) * sq(H, lh) / N_non_missing;
    if (alpha[1] != 0)        p += 0.5 * alpha[1] * gs(W, d.k, d.n) / N_non_missing;
    if (beta[1] != 0)         p += 0.5 * beta[1] * gs(H, d.k, d.m) / N_non_missing;
    if (alpha[2] != 0)        p += alpha[2] * sm(W, lw) / N_non_missing;
    if (beta[2] != 0)         p += beta[2] * sm(H, lh) / N_non_missing;
    return p;
}

// mse and the variable part of mkl for the current factors: src/nnmf.cpp:121-141 (Ahat = W.t()*H is evaluated
// column by column instead of being materialised; same arithmetic). Missing entries are skipped (:123-125).
void eval_errors(const double* A, const double* Wt, const double* H, Dims d, int nt, double N_non_missing,
                 double* mse, double* mkl_var)
{
    const int k = d.k; const int64_t n = d.n, m = d.m;
    double s2 = 0, skl = 0;
    #pragma omp parallel for num_threads(nt) schedule(static) reduction(+:s2, skl)
    for (int64_t j = 0; j < m; j++) {
        const double* h = H + (size_t)k * j;
        const double* Aj = A + (size_t)n * j;
        for (int64_t i = 0; i < n; i++) {
            const double a = Aj[i];
            if (!std::isfinite(a)) continue;
            const double* w = Wt + (size_t)k * i;
            double ah = 0;
            for (int c = 0; c < k; c++) ah += w[c] * h[c];
            const double r = a - ah;
            s2 += r * r;
            skl += -(a + TINY) * std::log(ah + TINY) + ah;
        }
    }
    *mse = s2 / N_non_missing;
    *mkl_var = skl / N_non_missing;
}

void set_err(char* err, size_t errlen, const char* msg)
{
    if (err && errlen) { std::snprintf(err, errlen, "%s", msg); }
}

}  // namespace

extern "C" {

void oracle_set_faithful_transpose(int on) { g_faithful_transpose = on; }

// Out-of-place transposition At (m x n) = A (n x m)': the copy the reference makes every outer iteration by binding
// A.t() to a const mat& (src/nnmf.cpp:117,131). Exposed so the CPU baseline can time it with the half-iterations.
void oracle_transpose(const double* A, int64_t n, int64_t m, double* At, int32_t n_threads)
{
    const int nt = resolve_threads(n_threads);
    constexpr int64_t B = 64;
    #pragma omp parallel for num_threads(nt) schedule(static) collapse(2)
    for (int64_t i0 = 0; i0 < n; i0 += B)
        for (int64_t j0 = 0; j0 < m; j0 += B) {
            const int64_t i1 = std::min(n, i0 + B), j1 = std::min(m, j0 + B);
            for (int64_t i = i0; i < i1; i++)
                for (int64_t j = j0; j < j1; j++) At[j + (size_t)m * i] = A[i + (size_t)n * j];
        }
}

// Synthetic workload of SURVEY.md §8(d) / BASELINE.md §4 on the host (the twin of the device generator the GPU arm uses, so
// the CPU reference arm of bench.py needs no GPU library): rows [row0, row0+nr) x columns [col0, col0+mc) of
//   A = u(base+1)(n_global x k) * u(base+2)(k x m) + noise * u(base+3, i + n_global*j),  NaN iff u(base+4, idx) < na_frac,
// with u(seed, idx) = (splitmix64(seed*0x9E3779B97F4A7C15 + idx) >> 11) * 2^-53. The inner product is one fused
// multiply-add chain in coordinate order, exactly as the device kernel forms it, so both generators agree bit for bit.
static inline double splitmix_u(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

void oracle_synth_block(double* A, int64_t n_global, int64_t row0, int64_t nr, int64_t col0, int64_t mc, int32_t k,
                        uint64_t base, double noise, double na_frac, int32_t n_threads)
{
    const int nt = resolve_threads(n_threads);
    std::vector<double> Wt((size_t)nr * k);          // [c][i] like the device copy: unit stride along i
    for (int c = 0; c < k; c++)
        for (int64_t i = 0; i < nr; i++) Wt[(size_t)c * nr + i] = splitmix_u(base + 1, (uint64_t)(row0 + i) + (uint64_t)n_global * (uint64_t)c);
    const double nanv = std::nan("");
    #pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t j = 0; j < mc; j++) {
        std::vector<double> h(k);
        for (int c = 0; c < k; c++) h[c] = splitmix_u(base + 2, (uint64_t)k * (uint64_t)(col0 + j) + (uint64_t)c);
        double* Aj = A + (size_t)nr * j;
        for (int64_t i = 0; i < nr; i++) Aj[i] = 0.0;
        for (int c = 0; c < k; c++) {
            const double* w = Wt.data() + (size_t)c * nr;
            const double hc = h[c];
            for (int64_t i = 0; i < nr; i++) Aj[i] = std::fma(w[i], hc, Aj[i]);
        }
        for (int64_t i = 0; i < nr; i++) {
            const uint64_t idx = (uint64_t)(row0 + i) + (uint64_t)n_global * (uint64_t)(col0 + j);
            double s = std::fma(noise, splitmix_u(base + 3, idx), Aj[i]);
            if (na_frac > 0.0 && splitmix_u(base + 4, idx) < na_frac) s = nanv;
            Aj[i] = s;
        }
    }
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// one half-iteration; same contract as nnlm_update (include/nnlm_b200.h)
// torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py calls this with the
// number of cores the process may run on so "all host cores" stays true under torchrun.
void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_update(double* H, const double* Wt, const double* A, const int32_t* mask, const double* beta,
                  int32_t k, int64_t n, int64_t m, uint32_t max_iter, double rel_tol, int32_t n_threads,
                  int32_t method, int32_t with_missing, int64_t* total_iter,
                  const void* /*opt*/, void* /*stats*/, char* err, size_t errlen)
{
    if (!H || !Wt || !A || !beta || k <= 0 || n < 0 || m < 0 || method < 1 || method > 4) {
        set_err(err, errlen, "oracle_update: bad argument"); return -1;
    }
    Dims d{k, n, m};
    if (with_missing < 0) {
        with_missing = 0;
        for (size_t e = 0; e < (size_t)n * m; e++) if (!std::isfinite(A[e])) { with_missing = 1; break; }
    }
    const int64_t t = with_missing ? update_missing(H, Wt, A, 1, n, mask, beta, d, max_iter, rel_tol, n_threads, method)
                                   : update_dense(H, Wt, A, 1, n, mask, beta, d, max_iter, rel_tol, n_threads, method);
    if (total_iter) *total_iter = t;
    return 0;
}

// src/nnmf.cpp:4-220; same contract as nnlm_nnmf (include/nnlm_b200.h)
int oracle_nnmf(const double* A, int64_t n, int64_t m, int32_t K,
                double* W, double* H, const int32_t* Wm, const int32_t* Hm,
                const double* alpha, const double* beta,
                uint32_t max_iter, double rel_tol, int32_t n_threads, int32_t verbose,
                uint32_t inner_max_iter, double inner_rel_tol, int32_t method, uint32_t trace,
                double* mse, double* mkl, double* target, double* avg_epoch, uint32_t err_cap,
                uint32_t* n_err, uint32_t* n_iter, int32_t* converged,
                int (*interrupt)(void*), void* interrupt_user,
                const void* /*opt*/, void* /*stats*/, char* err, size_t errlen)
{
    if (!A || !W || !H || !alpha || !beta || K <= 0 || n <= 0 || m <= 0 || method < 1 || method > 4) {
        set_err(err, errlen, "oracle_nnmf: bad argument"); return -1;
    }
    const int k = K;
    if (trace < 1) trace = 1;                                                   // :53
    const uint32_t err_len = (uint32_t)std::ceil((double)max_iter / (double)trace) + 1;   // :54
    if (err_cap < err_len || !mse || !mkl || !target || !avg_epoch) {
        set_err(err, errlen, "oracle_nnmf: error vectors too short"); return -1;
    }
    const int nt = resolve_threads(n_threads);
    const size_t NM = (size_t)n * m;

    // :64-73 missing detection and the constant part of the KL distance
    double N_non_missing = (double)NM;   // the reference holds this in an unsigned int (:51); 64-bit here
    bool any_missing = false;
    {
        double c = 0; int64_t cnt = 0;
        #pragma omp parallel for num_threads(nt) reduction(+:c, cnt)
        for (int64_t e = 0; e < (int64_t)NM; e++) {
            const double a = A[e];
            if (std::isfinite(a)) { c += (a + TINY) * std::log(a + TINY) - a; cnt++; }
        }
        any_missing = (size_t)cnt != NM;
        N_non_missing = (double)cnt;
        const double mkl_const = c / N_non_missing;
        for (uint32_t e = 0; e < err_len; e++) mkl[e] = mkl_const;
    }

    // :75-98 — W and the W mask are held transposed (k x n) for the whole run
    std::vector<double> Wt((size_t)k * n);
    for (int64_t i = 0; i < n; i++) for (int c = 0; c < k; c++) Wt[c + (size_t)k * i] = W[i + (size_t)n * c];
    std::vector<int32_t> Wmt;
    if (Wm) { Wmt.resize((size_t)k * n); for (int64_t i = 0; i < n; i++) for (int c = 0; c < k; c++) Wmt[c + (size_t)k * i] = Wm[i + (size_t)n * c]; }
    const int32_t* wmask = Wm ? Wmt.data() : nullptr;

    std::vector<double> At;   // explicit transposed copy for the faithful-timing variant
    auto w_half = [&]() -> int64_t {
        Dims d{k, m, n};   // roles swapped: solve for W (k x n) given "Wt" := H (k x m) and A.t() (m x n)
        if (g_faithful_transpose) {
            At.resize(NM);
            #pragma omp parallel for num_threads(nt) schedule(static)
            for (int64_t i = 0; i < n; i++) for (int64_t j = 0; j < m; j++) At[j + (size_t)m * i] = A[i + (size_t)n * j];
            return any_missing ? update_missing(Wt.data(), H, At.data(), 1, m, wmask, alpha, d, inner_max_iter, inner_rel_tol, n_threads, method)
                               : update_dense(Wt.data(), H, At.data(), 1, m, wmask, alpha, d, inner_max_iter, inner_rel_tol, n_threads, method);
        }
        // A.t()(jj, ii) = A(ii, jj): element (row=jj over m, col=ii over n) sits at A[ii + n*jj]
        return any_missing ? update_missing(Wt.data(), H, A, n, 1, wmask, alpha, d, inner_max_iter, inner_rel_tol, n_threads, method)
                           : update_dense(Wt.data(), H, A, n, 1, wmask, alpha, d, inner_max_iter, inner_rel_tol, n_threads, method);
    };
    auto h_half = [&]() -> int64_t {
        Dims d{k, n, m};
        return any_missing ? update_missing(H, Wt.data(), A, 1, n, Hm, beta, d, inner_max_iter, inner_rel_tol, n_threads, method)
                           : update_dense(H, Wt.data(), A, 1, n, Hm, beta, d, inner_max_iter, inner_rel_tol, n_threads, method);
    };

    double rel_err = rel_tol + 1;      // :62
    double terr_last = 1e99;           // :63
    int64_t total_raw_iter = 0;
    uint32_t i = 0, i_e = 0;
    const Dims dfull{k, n, m};

    auto record = [&]() {              // :121-160 and the tail :164-192 share this bookkeeping
        double e_mse, e_kl;
        eval_errors(A, Wt.data(), H, dfull, nt, N_non_missing, &e_mse, &e_kl);
        mse[i_e] = e_mse;
        mkl[i_e] += e_kl;
        avg_epoch[i_e] = (double)total_raw_iter / (double)(n + m);
        target[i_e] = (method < 3) ? 0.5 * mse[i_e] : mkl[i_e];
        target[i_e] += penalty(Wt.data(), H, dfull, N_non_missing, alpha, beta);
        rel_err = 2 * (terr_last - target[i_e]) / (terr_last + target[i_e] + TINY);
        terr_last = target[i_e];
        if (verbose == 2) std::printf("%10u | %10.4f | %10.4f | %10.4f | %10.g\n", i + 1, mse[i_e], mkl[i_e], target[i_e], rel_err);
        total_raw_iter = 0;
        ++i_e;
    };

    for (; i < max_iter && std::fabs(rel_err) > rel_tol; i++) {                // :109
        if (interrupt && interrupt(interrupt_user)) { set_err(err, errlen, "interrupted"); return -5; }   // :111
        total_raw_iter += w_half();                                             // :117 / :131
        total_raw_iter += h_half();                                             // :119 / :133
        if (i % trace == 0) record();                                           // :143-160
    }
    if ((uint32_t)(i - 1) % trace != 0) record();                               // :164 (unsigned wrap kept)

    if (n_err) *n_err = i_e;                                                    // :200-206
    if (n_iter) *n_iter = i;
    if (converged) *converged = !(rel_err > rel_tol);                           // :208
    for (int64_t ii = 0; ii < n; ii++) for (int c = 0; c < k; c++) W[ii + (size_t)n * c] = Wt[c + (size_t)k * ii];   // W.t(), :212
    return 0;
}

// src/nnlm.cpp:36-52; same contract as nnlm_nnlm (include/nnlm_b200.h)
int oracle_nnlm(const double* x, const double* y, int64_t n, int64_t p, int64_t q,
                double* coef, const int32_t* mask, const double* alpha,
                uint32_t max_iter, double rel_tol, int32_t n_threads, int32_t method, int64_t* n_iteration,
                const void* /*opt*/, void* /*stats*/, char* err, size_t errlen)
{
    if (!x || !y || !coef || !alpha || n <= 0 || p <= 0 || q <= 0 || p > INT32_MAX || method < 1 || method > 4) {
        set_err(err, errlen, "oracle_nnlm: bad argument"); return -1;
    }
    std::vector<double> xt((size_t)p * n);                                      // x.t()
    for (int64_t i = 0; i < n; i++) for (int64_t c = 0; c < p; c++) xt[c + (size_t)p * i] = x[i + (size_t)n * c];
    bool any_missing = false;
    for (size_t e = 0; e < (size_t)n * q; e++) if (!std::isfinite(y[e])) { any_missing = true; break; }
    Dims d{(int)p, n, q};
    const int64_t t = any_missing ? update_missing(coef, xt.data(), y, 1, n, mask, alpha, d, max_iter, rel_tol, n_threads, method)
                                  : update_dense(coef, xt.data(), y, 1, n, mask, alpha, d, max_iter, rel_tol, n_threads, method);
    if (n_iteration) *n_iteration = t;
    return 0;
}

}  // extern "C"
