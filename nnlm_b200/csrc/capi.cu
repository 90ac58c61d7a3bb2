// capi.cu — the C ABI of include/nnlm_b200.h: argument checking, the outer ANLS loop of c_nnmf (src/nnmf.cpp:48-220),
// c_nnlm (src/nnlm.cpp:36-52), a single update() (src/update_with_missing.cpp), and the device-resident session used by
// the benchmark. All arithmetic happens in the kernels driven by Engine; this file is bookkeeping.
#include <chrono>
#include <cmath>
#include <memory>
#include <vector>

#include "engine.cuh"

struct nnlm_comm;

using namespace nnlm;

namespace {

void set_err(char* err, size_t errlen, const char* msg)
{
    if (err && errlen) std::snprintf(err, errlen, "%s", msg);
}

int device_count_noexcept()
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

template <typename F>
int guarded(char* err, size_t errlen, F&& f)
{
    if (err && errlen) err[0] = 0;
    if (device_count_noexcept() <= 0) {
        set_err(err, errlen, "nnlm_b200: no CUDA device available (this library has no CPU fallback)");
        return NNLM_E_NO_DEVICE;
    }
    try {
        return f();
    } catch (const Error& e) {
        set_err(err, errlen, e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        set_err(err, errlen, "nnlm_b200: host allocation failed");
        return NNLM_E_NOMEM;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return NNLM_E_CUDA;
    }
}

void fill_stats(nnlm_stats* st, const Engine& e, uint64_t launches0)
{
    if (!st) return;
    st->launches = launch_counter().load() - launches0;
    st->h2d_bytes = e.h2d_bytes;
    st->d2h_bytes = e.d2h_bytes;
    st->precision_used = e.precision_used();
    st->cross_ms = e.timer.ms[KernelTimer::CROSS];
    st->solve_ms = e.timer.ms[KernelTimer::SOLVE];
    st->error_ms = e.timer.ms[KernelTimer::ERROR];
    st->gram_ms = e.timer.ms[KernelTimer::GRAM];
    st->comm_ms = e.timer.ms[KernelTimer::COMM];
    st->comm_bytes = e.comm_bytes;
    st->cross_launches = e.timer.count[KernelTimer::CROSS];
    st->solve_launches = e.timer.count[KernelTimer::SOLVE];
}

float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

struct Events {
    cudaEvent_t e[4];
    Events() { for (auto& x : e) cudaEventCreate(&x); }
    ~Events() { for (auto& x : e) cudaEventDestroy(x); }
};

// src/nnmf.cpp:224-240 from the factor statistics (sum X^2, sum X, accu(X X'))
double penalty_from_stats(const ErrorTerms& t, const double* alpha, const double* beta, double N)
{
    double p = 0;
    if (alpha[0] != alpha[1]) p += 0.5 * (alpha[0] - alpha[1]) * t.w_stats[0] / N;
    if (beta[0] != beta[1])   p += 0.5 * (beta[0] - beta[1]) * t.h_stats[0] / N;
    if (alpha[1] != 0)        p += 0.5 * alpha[1] * t.w_stats[2] / N;
    if (beta[1] != 0)         p += 0.5 * beta[1] * t.h_stats[2] / N;
    if (alpha[2] != 0)        p += alpha[2] * t.w_stats[1] / N;
    if (beta[2] != 0)         p += beta[2] * t.h_stats[1] / N;
    return p;
}

int precision_of(const nnlm_options* opt, int64_t n, int64_t m)
{
    int p = opt ? opt->precision : NNLM_PREC_AUTO;
    if (p == NNLM_PREC_AUTO) p = ((double)n * (double)m >= 4.0e6) ? NNLM_PREC_FAST : NNLM_PREC_EXACT;
    return p;
}

}  // namespace

struct nnlm_session {
    std::unique_ptr<Engine> eng;
    uint64_t launches0 = 0;
    double loop_ms = 0;
    double upload_ms = 0;
};

extern "C" {
#pragma GCC visibility push(default)

int nnlm_abi_version(void) { return NNLM_B200_ABI_VERSION; }

int nnlm_device_count(char* name, size_t namelen)
{
    const int c = device_count_noexcept();
    if (name && namelen) {
        name[0] = 0;
        if (c > 0) {
            cudaDeviceProp p;
            if (cudaGetDeviceProperties(&p, 0) == cudaSuccess) std::snprintf(name, namelen, "%s", p.name);
        }
    }
    return c;
}

int nnlm_nnmf(const double* A, int64_t n, int64_t m, int32_t K,
              double* W, double* H, const int32_t* Wm, const int32_t* Hm,
              const double* alpha, const double* beta,
              uint32_t max_iter, double rel_tol, int32_t /*n_threads*/, int32_t verbose,
              uint32_t inner_max_iter, double inner_rel_tol, int32_t method, uint32_t trace,
              double* mse, double* mkl, double* target, double* avg_epoch, uint32_t err_cap,
              uint32_t* n_err, uint32_t* n_iter, int32_t* converged,
              nnlm_interrupt_fn interrupt, void* interrupt_user,
              const nnlm_options* opt, nnlm_stats* stats,
              char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(A && W && H && alpha && beta, "nnlm_nnmf: NULL argument");
        NNLM_REQUIRE(n > 0 && m > 0 && K > 0, "nnlm_nnmf: dimensions must be positive");
        NNLM_REQUIRE(method >= 1 && method <= 4, "nnlm_nnmf: method code must be 1..4");
        if (trace < 1) trace = 1;                                                            // src/nnmf.cpp:53
        const uint32_t err_len = (uint32_t)std::ceil((double)max_iter / (double)trace) + 1;  // :54
        NNLM_REQUIRE(mse && mkl && target && avg_epoch && err_cap >= err_len,
                     "nnlm_nnmf: error vectors must hold ceil(max_iter/trace)+1 entries");
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        Events ev;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms_since = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count(); };
        const auto h0 = now();

        Engine eng(n, m, K, method, precision_of(opt, n, m), opt ? opt->device : -1);
        eng.timer.enable(opt && opt->verbose_timing);
        cudaStream_t st = eng.stream();
        cudaEventRecord(ev.e[0], st);
        eng.upload_A(A);                                        // + missing detection and KL constant, :64-73
        eng.set_factors(W, H);                                  // :82-98 (explicit init; the shim draws the default one)
        eng.set_masks(Wm, Hm);
        eng.set_penalties(alpha, beta);
        eng.set_inner(inner_max_iter, inner_rel_tol);
        cudaEventRecord(ev.e[1], st);
        eng.sync();
        const auto h1 = now();

        const double N = (double)((int64_t)n * m - eng.n_missing());   // N_non_missing, :51,68
        const double mkl_const = eng.kl_const_sum() / N;               // :70-73
        for (uint32_t e = 0; e < err_len; e++) mkl[e] = mkl_const;

        double rel_err = rel_tol + 1;     // :62
        double terr_last = 1e99;          // :63
        uint32_t i = 0, i_e = 0;
        uint64_t total_raw_iter = 0;

        auto record = [&]() {             // :121-160 and the tail :164-192
            ErrorTerms t;
            eng.errors(&t);
            total_raw_iter += eng.take_sweeps();
            mse[i_e] = t.sum_sq / N;
            mkl[i_e] += t.sum_kl / N;
            avg_epoch[i_e] = (double)total_raw_iter / (double)(n + m);
            target[i_e] = (method < 3) ? 0.5 * mse[i_e] : mkl[i_e];
            target[i_e] += penalty_from_stats(t, alpha, beta, N);
            rel_err = 2 * (terr_last - target[i_e]) / (terr_last + target[i_e] + TINY_NUM);
            terr_last = target[i_e];
            if (verbose == 2)
                std::printf("%10u | %10.4f | %10.4f | %10.4f | %10.g\n", i + 1, mse[i_e], mkl[i_e], target[i_e], rel_err);
            total_raw_iter = 0;
            ++i_e;
        };

        if (verbose == 2) {
            std::printf("\n%10s | %10s | %10s | %10s | %10s\n", "Iteration", "MSE", "MKL", "Target", "Rel. Err.");
            std::printf("--------------------------------------------------------------\n");
        }
        for (; i < max_iter && std::fabs(rel_err) > rel_tol; i++) {                          // :109
            if (interrupt && interrupt(interrupt_user)) {                                    // :111
                eng.sync();
                set_err(err, errlen, "nnlm_nnmf: interrupted");
                return NNLM_E_INTERRUPT;
            }
            eng.half_w();                                                                    // :117 / :131
            eng.half_h();                                                                    // :119 / :133
            if (i % trace == 0) record();                                                    // :143
        }
        if ((uint32_t)(i - 1) % trace != 0) record();                                        // :164
        if (verbose == 2) {
            std::printf("--------------------------------------------------------------\n");
            std::printf("%10s | %10s | %10s | %10s | %10s\n\n", "Iteration", "MSE", "MKL", "Target", "Rel. Err.");
        }
        cudaEventRecord(ev.e[2], st);
        eng.sync();
        const auto h2 = now();
        eng.get_factors(W, H);                                                               // :211-213
        cudaEventRecord(ev.e[3], st);
        eng.sync();
        const auto h3 = now();

        if (n_err) *n_err = i_e;                                                             // :200-206
        if (n_iter) *n_iter = i;                                                             // :218
        if (converged) *converged = !(rel_err > rel_tol);                                    // :208
        if (stats) {
            fill_stats(stats, eng, launches0);
            stats->upload_ms = elapsed(ev.e[0], ev.e[1]);
            stats->loop_ms = elapsed(ev.e[1], ev.e[2]);
            stats->download_ms = elapsed(ev.e[2], ev.e[3]);
            stats->host_setup_ms = ms_since(h0, h1);
            stats->host_loop_ms = ms_since(h1, h2);
            stats->host_finish_ms = ms_since(h2, h3);
        }
        return NNLM_OK;
    });
}

int nnlm_update(double* H, const double* Wt, const double* A, const int32_t* mask, const double* beta,
                int32_t k, int64_t n, int64_t m,
                uint32_t max_iter, double rel_tol, int32_t /*n_threads*/, int32_t method, int32_t with_missing,
                int64_t* total_iter,
                const nnlm_options* opt, nnlm_stats* stats,
                char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(H && Wt && A && beta, "nnlm_update: NULL argument");
        NNLM_REQUIRE(n > 0 && m > 0 && k > 0, "nnlm_update: dimensions must be positive");
        NNLM_REQUIRE(method >= 1 && method <= 4, "nnlm_update: method code must be 1..4");
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        Engine eng(n, m, k, method, precision_of(opt, n, m), opt ? opt->device : -1, /*both_sides=*/false);
        eng.set_missing_mode(with_missing);
        eng.upload_A(A);
        eng.set_factors_t(Wt, H);
        eng.set_masks(nullptr, mask);
        eng.set_penalties(nullptr, beta);
        eng.set_inner(max_iter, rel_tol);
        eng.half_h();
        const uint64_t t = eng.take_sweeps();
        eng.get_H(H);
        if (total_iter) *total_iter = (int64_t)t;
        fill_stats(stats, eng, launches0);
        return NNLM_OK;
    });
}

int nnlm_nnlm(const double* x, const double* y, int64_t n, int64_t p, int64_t q,
              double* coef, const int32_t* mask, const double* alpha,
              uint32_t max_iter, double rel_tol, int32_t /*n_threads*/, int32_t method,
              int64_t* n_iteration,
              const nnlm_options* opt, nnlm_stats* stats,
              char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(x && y && coef && alpha, "nnlm_nnlm: NULL argument");
        NNLM_REQUIRE(n > 0 && p > 0 && q > 0 && p <= INT32_MAX, "nnlm_nnlm: dimensions must be positive");
        NNLM_REQUIRE(method >= 1 && method <= 4, "nnlm_nnlm: method code must be 1..4");
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        // update(beta, x.t(), y, mask, alpha, ...)  (src/nnlm.cpp:44-47): the engine's "W" is x (n x p), its "A" is y
        Engine eng(n, q, (int)p, method, precision_of(opt, n, q), opt ? opt->device : -1, /*both_sides=*/false);
        eng.upload_A(y);
        eng.set_factors(x, coef);                                // x.t() is formed on the device
        eng.set_masks(nullptr, mask);
        eng.set_penalties(nullptr, alpha);
        eng.set_inner(max_iter, rel_tol);
        eng.half_h();
        const uint64_t t = eng.take_sweeps();
        eng.get_H(coef);
        if (n_iteration) *n_iteration = (int64_t)t;
        fill_stats(stats, eng, launches0);
        return NNLM_OK;
    });
}

// diagnostic entry: the cross-product Q = Wt * A alone (the contraction the reference forms per column,
// src/update_with_missing.cpp:39), through the same kernels a half-iteration uses
int nnlm_cross(const double* Wt, const double* A, int32_t k, int64_t n, int64_t m, double* Q,
               const nnlm_options* opt, nnlm_stats* stats, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(Wt && A && Q && k > 0 && n > 0 && m > 0, "nnlm_cross: bad argument");
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        Engine eng(n, m, k, NNLM_SCD_MSE, precision_of(opt, n, m), opt ? opt->device : -1, /*both_sides=*/false);
        eng.set_missing_mode(0);
        eng.upload_A(A);
        std::vector<double> H0((size_t)k * m, 0.0);
        eng.set_factors_t(Wt, H0.data());
        eng.cross_only(Q);
        fill_stats(stats, eng, launches0);
        return NNLM_OK;
    });
}

// ---- device-resident session ---------------------------------------------------------------------------------------

int nnlm_session_create(nnlm_session** out, const double* A, int64_t n, int64_t m, int32_t K,
                        const int32_t* Wm, const int32_t* Hm,
                        const double* alpha, const double* beta,
                        uint32_t inner_max_iter, double inner_rel_tol, int32_t method,
                        const nnlm_options* opt, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(out && A, "nnlm_session_create: NULL argument");
        std::unique_ptr<nnlm_session> s(new nnlm_session);
        s->launches0 = launch_counter().load();
        s->eng.reset(new Engine(n, m, K, method, precision_of(opt, n, m), opt ? opt->device : -1));
        s->eng->timer.enable(opt && opt->verbose_timing);
        Events ev;
        cudaEventRecord(ev.e[0], s->eng->stream());
        s->eng->upload_A(A);
        cudaEventRecord(ev.e[1], s->eng->stream());
        s->eng->sync();
        s->upload_ms = elapsed(ev.e[0], ev.e[1]);
        s->eng->set_masks(Wm, Hm);
        s->eng->set_penalties(alpha, beta);
        s->eng->set_inner(inner_max_iter, inner_rel_tol);
        *out = s.release();
        return NNLM_OK;
    });
}

int nnlm_session_create_synthetic(nnlm_session** out, int64_t n, int64_t m, int32_t K, uint64_t seed_base, double noise,
                                  double na_frac, const double* alpha, const double* beta,
                                  uint32_t inner_max_iter, double inner_rel_tol, int32_t method,
                                  const nnlm_options* opt, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(out, "nnlm_session_create_synthetic: NULL argument");
        std::unique_ptr<nnlm_session> s(new nnlm_session);
        s->launches0 = launch_counter().load();
        Comm* comm = (opt && opt->comm) ? nnlm_comm_get(static_cast<nnlm_comm*>(opt->comm)) : nullptr;
        s->eng.reset(new Engine(n, m, K, method, precision_of(opt, n, m), opt ? opt->device : -1, true, comm));
        s->eng->timer.enable(opt && opt->verbose_timing);
        Engine& e = *s->eng;
        if (!comm) {
            DevBuf<double> dA((size_t)n * m);
            launch_synth(dA.p, n, m, K, 0, seed_base, noise, na_frac, e.stream());
            e.ingest_shards(dA.p, dA.p);
        } else {
            // every rank generates exactly its column shard and its row shard of the global matrix
            DevBuf<double> dC((size_t)n * std::max<int64_t>(e.cols_local(), 1)), dR((size_t)std::max<int64_t>(e.rows_local(), 1) * m);
            launch_synth_block(dC.p, n, 0, n, e.col0(), e.cols_local(), K, seed_base, noise, na_frac, e.stream());
            launch_synth_block(dR.p, n, e.row0(), e.rows_local(), 0, m, K, seed_base, noise, na_frac, e.stream());
            e.ingest_shards(dC.p, dR.p);
        }
        e.set_penalties(alpha, beta);
        e.set_inner(inner_max_iter, inner_rel_tol);
        *out = s.release();
        return NNLM_OK;
    });
}

int nnlm_session_create_sharded(nnlm_session** out, const double* Acol, const double* Arow, int64_t n, int64_t m, int32_t K,
                                const int32_t* Wm, const int32_t* Hm, const double* alpha, const double* beta,
                                uint32_t inner_max_iter, double inner_rel_tol, int32_t method,
                                const nnlm_options* opt, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(out && Acol && Arow, "nnlm_session_create_sharded: NULL argument");
        std::unique_ptr<nnlm_session> s(new nnlm_session);
        s->launches0 = launch_counter().load();
        Comm* comm = (opt && opt->comm) ? nnlm_comm_get(static_cast<nnlm_comm*>(opt->comm)) : nullptr;
        s->eng.reset(new Engine(n, m, K, method, precision_of(opt, n, m), opt ? opt->device : -1, true, comm));
        s->eng->timer.enable(opt && opt->verbose_timing);
        Engine& e = *s->eng;
        Events ev;
        cudaEventRecord(ev.e[0], e.stream());
        const size_t cc = (size_t)n * e.cols_local(), cr = (size_t)e.rows_local() * m;
        DevBuf<double> dC(std::max<size_t>(cc, 1)), dR(std::max<size_t>(cr, 1));
        NNLM_CUDA_CHECK(cudaMemcpyAsync(dC.p, Acol, cc * sizeof(double), cudaMemcpyHostToDevice, e.stream()));
        NNLM_CUDA_CHECK(cudaMemcpyAsync(dR.p, Arow, cr * sizeof(double), cudaMemcpyHostToDevice, e.stream()));
        e.h2d_bytes += (cc + cr) * sizeof(double);
        e.ingest_shards(dC.p, dR.p);
        cudaEventRecord(ev.e[1], e.stream());
        e.sync();
        s->upload_ms = elapsed(ev.e[0], ev.e[1]);
        e.set_masks(Wm, Hm);
        e.set_penalties(alpha, beta);
        e.set_inner(inner_max_iter, inner_rel_tol);
        *out = s.release();
        return NNLM_OK;
    });
}

int nnlm_synth_block(double* A, int64_t n_global, int64_t row0, int64_t nr, int64_t col0, int64_t mc, int32_t k,
                     uint64_t seed_base, double noise, double na_frac, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(A && n_global > 0 && nr > 0 && mc > 0 && k > 0, "nnlm_synth_block: bad argument");
        DevBuf<double> dA((size_t)nr * mc);
        launch_synth_block(dA.p, n_global, row0, nr, col0, mc, k, seed_base, noise, na_frac, 0);
        NNLM_CUDA_CHECK(cudaMemcpy(A, dA.p, dA.bytes(), cudaMemcpyDeviceToHost));
        return NNLM_OK;
    });
}

int nnlm_synth_matrix(double* A, int64_t n, int64_t m, int32_t k, int64_t col0, uint64_t seed_base, double noise,
                      double na_frac, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(A && n > 0 && m > 0 && k > 0, "nnlm_synth_matrix: bad argument");
        DevBuf<double> dA((size_t)n * m);
        launch_synth(dA.p, n, m, k, col0, seed_base, noise, na_frac, 0);
        NNLM_CUDA_CHECK(cudaMemcpy(A, dA.p, dA.bytes(), cudaMemcpyDeviceToHost));
        return NNLM_OK;
    });
}

int nnlm_session_set_factors(nnlm_session* s, const double* W, const double* H, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(s && W && H, "nnlm_session_set_factors: NULL argument");
        s->eng->set_factors(W, H);
        return NNLM_OK;
    });
}

int nnlm_session_get_factors(nnlm_session* s, double* W, double* H, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(s && W && H, "nnlm_session_get_factors: NULL argument");
        s->eng->get_factors(W, H);
        return NNLM_OK;
    });
}

int nnlm_session_run(nnlm_session* s, uint32_t iters, double* device_ms, int64_t* total_sweeps, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(s, "nnlm_session_run: NULL session");
        Events ev;
        cudaStream_t st = s->eng->stream();
        s->eng->sync();
        cudaEventRecord(ev.e[0], st);
        for (uint32_t i = 0; i < iters; i++) { s->eng->half_w(); s->eng->half_h(); }          // src/nnmf.cpp:109-133
        cudaEventRecord(ev.e[1], st);
        s->eng->sync();
        const double ms = elapsed(ev.e[0], ev.e[1]);
        s->loop_ms += ms;
        if (device_ms) *device_ms = ms;
        const uint64_t t = s->eng->take_sweeps();
        if (total_sweeps) *total_sweeps = (int64_t)t;
        return NNLM_OK;
    });
}

int nnlm_session_error(nnlm_session* s, double* mse, double* mkl, double* target, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(s, "nnlm_session_error: NULL session");
        Engine& e = *s->eng;
        ErrorTerms t;
        e.errors(&t);
        const double N = (double)(e.n() * e.m() - e.n_missing());
        const double v_mse = t.sum_sq / N;
        const double v_mkl = e.kl_const_sum() / N + t.sum_kl / N;
        if (mse) *mse = v_mse;
        if (mkl) *mkl = v_mkl;
        if (target) *target = (e.method() < 3) ? 0.5 * v_mse : v_mkl;   // penalties: see nnlm_nnmf
        return NNLM_OK;
    });
}

int nnlm_session_stats(nnlm_session* s, nnlm_stats* stats)
{
    if (!s || !stats) return NNLM_E_ARG;
    std::memset(stats, 0, sizeof *stats);
    fill_stats(stats, *s->eng, s->launches0);
    stats->loop_ms = s->loop_ms;
    stats->upload_ms = s->upload_ms;
    return NNLM_OK;
}

int nnlm_session_reset_stats(nnlm_session* s)
{
    if (!s) return NNLM_E_ARG;
    s->eng->timer.reset();
    s->launches0 = launch_counter().load();
    s->loop_ms = 0;
    return NNLM_OK;
}

void nnlm_session_destroy(nnlm_session* s) { delete s; }

// ---- bit-exact NA mask ----------------------------------------------------------------------------------------------

int nnlm_na_mask(const double* A, int64_t n, int64_t m, uint32_t* bits, int64_t* col_missing, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(A && bits && n > 0 && m > 0, "nnlm_na_mask: bad argument");
        const size_t cnt = (size_t)n * m, words = (cnt + 31) / 32;
        DevBuf<double> dA(cnt);
        DevBuf<uint32_t> dB(words);
        DevBuf<int64_t> dC;
        if (col_missing) dC.alloc(m);
        NNLM_CUDA_CHECK(cudaMemcpy(dA.p, A, cnt * sizeof(double), cudaMemcpyHostToDevice));
        launch_na_bits(dA.p, n, m, dB.p, col_missing ? dC.p : nullptr, 0);
        NNLM_CUDA_CHECK(cudaMemcpy(bits, dB.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (col_missing) NNLM_CUDA_CHECK(cudaMemcpy(col_missing, dC.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost));
        return NNLM_OK;
    });
}

#pragma GCC visibility pop
}  // extern "C"
