// capi.cu — the C ABI of include/nnlm_b200.h: argument checking, the outer ANLS loop of c_nnmf (src/nnmf.cpp:48-220),
// c_nnlm (src/nnlm.cpp:36-52), a single update() (src/update_with_missing.cpp), and the device-resident session used by
// the benchmark. All arithmetic happens in the kernels driven by Engine; this file is bookkeeping.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
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

float elapsed(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    return cudaEventElapsedTime(&ms, a, b) == cudaSuccess ? ms : 0.0f;
}

struct Events {
    cudaEvent_t e[4];
    Events() { for (auto& x : e) NNLM_CUDA_CHECK(cudaEventCreate(&x)); }       // on the CURRENT device: create under ScopedDevice
    ~Events() { for (auto& x : e) cudaEventDestroy(x); }
    void record(int i, cudaStream_t st) { NNLM_CUDA_CHECK(cudaEventRecord(e[i], st)); }
};

// selects `dev` for the lifetime of the object and restores the caller's device afterwards
struct ScopedDevice {
    int prev = -1;
    explicit ScopedDevice(int dev) { cudaGetDevice(&prev); if (dev >= 0 && dev != prev) NNLM_CUDA_CHECK(cudaSetDevice(dev)); }
    ~ScopedDevice() { if (prev >= 0) cudaSetDevice(prev); }
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

// AUTO: whole factorisations (nnlm_nnmf, sessions) leave the choice to the engine, which knows the size and, after the
// ingest pass, the data (Engine::ingest_shards). The single-update entry points (nnlm_nnlm, nnlm_update) are called with
// tolerances down to 1e-12 and stay in fp64 unless FAST is asked for explicitly.
int precision_of(const nnlm_options* opt, int64_t /*n*/, int64_t /*m*/, bool factorisation)
{
    const int p = opt ? opt->precision : NNLM_PREC_AUTO;
    if (p == NNLM_PREC_AUTO) return factorisation ? NNLM_PREC_AUTO : NNLM_PREC_EXACT;
    return p;
}

int device_of(const nnlm_options* opt)
{
    int dev = opt ? opt->device : -1;
    if (dev < 0) NNLM_CUDA_CHECK(cudaGetDevice(&dev));
    return dev;
}

int gpus_requested(const nnlm_options* opt)
{
    int g = opt ? opt->n_gpus : 0;
    if (g <= 0) { const char* e = std::getenv("NNLM_B200_GPUS"); g = e ? std::atoi(e) : 1; }
    return g < 1 ? 1 : g;
}

// sense-reversing spin barrier for the worker threads of one multi-GPU call (at most 8 threads, microseconds apart)
class HostBarrier {
public:
    explicit HostBarrier(int n) : n_(n) {}
    void wait()
    {
        const unsigned gen = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        } else {
            while (gen_.load(std::memory_order_acquire) == gen) std::this_thread::yield();
        }
    }
private:
    const int n_;
    std::atomic<int> count_{0};
    std::atomic<unsigned> gen_{0};
};

struct NnmfCall {              // the arguments of c_nnmf (src/nnmf.cpp:4-9), shared read-only by every rank of the call
    const double* A; int64_t n, m; int32_t K;
    double* W; double* H; const int32_t* Wm; const int32_t* Hm;
    const double* alpha; const double* beta;
    uint32_t max_iter; double rel_tol; int32_t verbose; uint32_t inner_max_iter; double inner_rel_tol; int32_t method; uint32_t trace;
    double *mse, *mkl, *target, *avg_epoch;
    uint32_t *n_err, *n_iter; int32_t* converged;
    nnlm_interrupt_fn interrupt; void* interrupt_user;
    const nnlm_options* opt; nnlm_stats* stats;
    int precision;
};

struct Team {                  // shared state of the ranks of one multi-GPU call
    explicit Team(int n) : bar(n) {}
    HostBarrier bar;
    std::atomic<int> failed{0};
    std::atomic<int> stop{0};
};

void say(const NnmfCall& c, const char* fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c.opt && c.opt->print) c.opt->print(c.opt->print_user, buf);
    else std::fputs(buf, stdout);
}

// The outer ANLS loop of c_nnmf (src/nnmf.cpp:48-220) on one engine. With a team every rank runs the same loop on its
// shard: all loop decisions derive from all-reduced sums, which are bit-identical on every rank, so the ranks stay in step
// without host synchronisation; only the root (rank 0, the calling thread) polls the interrupt callback, prints and writes
// the outputs. Returns NNLM_OK or NNLM_E_INTERRUPT.
int anls_loop(Engine& eng, const NnmfCall& c, bool root, Team* team, std::chrono::steady_clock::time_point t_entry,
              uint64_t launches0)
{
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count(); };
    const int64_t n = c.n, m = c.m;
    const uint32_t trace = c.trace < 1 ? 1 : c.trace;                                        // src/nnmf.cpp:53
    const uint32_t err_len = (uint32_t)std::ceil((double)c.max_iter / (double)trace) + 1;    // :54
    const auto h1 = now();
    cudaStream_t st = eng.stream();
    Events ev;
    ev.record(1, st);

    const double N = (double)((int64_t)n * m - eng.n_missing());   // N_non_missing, :51,68
    const double mkl_const = eng.kl_const_sum() / N;               // :70-73
    std::vector<double> l_mse(err_len), l_mkl(err_len, mkl_const), l_tgt(err_len), l_ep(err_len);   // every rank keeps its own copy

    double rel_err = c.rel_tol + 1;   // :62
    double terr_last = 1e99;          // :63
    uint32_t i = 0, i_e = 0;
    uint64_t total_raw_iter = 0;
    // mkl_trace = 1: the square-loss methods need only the MSE for their target error; it comes from the Gram identity
    // (Engine::errors) and the KL distance is evaluated once, for the final record
    const bool lazy_kl = c.opt && c.opt->mkl_trace == 1 && c.method < 3;
    bool used_identity = false;

    auto record = [&]() {             // :121-160 and the tail :164-192
        ErrorTerms t;
        eng.errors(&t, !lazy_kl);
        used_identity = used_identity || eng.last_mse_from_identity();
        total_raw_iter += eng.take_sweeps();
        l_mse[i_e] = t.sum_sq / N;
        l_mkl[i_e] += t.sum_kl / N;
        l_ep[i_e] = (double)total_raw_iter / (double)(n + m);
        l_tgt[i_e] = (c.method < 3) ? 0.5 * l_mse[i_e] : l_mkl[i_e];
        l_tgt[i_e] += penalty_from_stats(t, c.alpha, c.beta, N);
        rel_err = 2 * (terr_last - l_tgt[i_e]) / (terr_last + l_tgt[i_e] + TINY_NUM);
        terr_last = l_tgt[i_e];
        if (c.verbose == 2 && root)
            say(c, "%10u | %10.4f | %10.4f | %10.4f | %10.g\n", i + 1, l_mse[i_e], l_mkl[i_e], l_tgt[i_e], rel_err);
        total_raw_iter = 0;
        ++i_e;
    };

    if (c.verbose == 2 && root) {
        say(c, "\n%10s | %10s | %10s | %10s | %10s\n", "Iteration", "MSE", "MKL", "Target", "Rel. Err.");
        say(c, "--------------------------------------------------------------\n");
    }
    bool interrupted = false;
    for (; i < c.max_iter && std::fabs(rel_err) > c.rel_tol; i++) {                          // :109
        if (c.interrupt) {                                                                   // :111
            if (root && c.interrupt(c.interrupt_user)) { if (team) team->stop.store(1); else interrupted = true; }
            if (team) { team->bar.wait(); interrupted = team->stop.load() != 0; team->bar.wait(); }
            if (interrupted) break;
        }
        eng.half_w();                                                                        // :117 / :131
        eng.half_h();                                                                        // :119 / :133
        if (i % trace == 0) record();                                                        // :143
    }
    if (interrupted) { eng.sync(); return NNLM_E_INTERRUPT; }
    if ((uint32_t)(i - 1) % trace != 0) record();                                            // :164
    if (lazy_kl && i_e > 0 && used_identity) {           // the KL distance of the final factors (one fused pass over A)
        ErrorTerms t;
        eng.errors(&t, true);
        l_mkl[i_e - 1] = mkl_const + t.sum_kl / N;
    }
    if (c.verbose == 2 && root) {
        say(c, "--------------------------------------------------------------\n");
        say(c, "%10s | %10s | %10s | %10s | %10s\n\n", "Iteration", "MSE", "MKL", "Target", "Rel. Err.");
    }
    ev.record(2, st);
    eng.sync();
    const auto h2 = now();
    if (root) {
        eng.get_factors(c.W, c.H);                                                           // :211-213
        ev.record(3, st);
        eng.sync();
        const auto h3 = now();
        for (uint32_t e = 0; e < i_e; e++) { c.mse[e] = l_mse[e]; c.mkl[e] = l_mkl[e]; c.target[e] = l_tgt[e]; c.avg_epoch[e] = l_ep[e]; }
        if (c.n_err) *c.n_err = i_e;                                                         // :200-206
        if (c.n_iter) *c.n_iter = i;                                                         // :218
        if (c.converged) *c.converged = !(rel_err > c.rel_tol);                              // :208
        if (c.stats) {
            fill_stats(c.stats, eng, launches0);
            c.stats->loop_ms = elapsed(ev.e[1], ev.e[2]);
            c.stats->download_ms = elapsed(ev.e[2], ev.e[3]);
            c.stats->host_setup_ms = ms_since(t_entry, h1);
            c.stats->host_loop_ms = ms_since(h1, h2);
            c.stats->host_finish_ms = ms_since(h2, h3);
            c.stats->mse_from_identity = used_identity ? 1 : 0;
        }
    }
    return NNLM_OK;
}

// one GPU: upload the whole matrix, run the loop
int nnmf_single(const NnmfCall& c, std::chrono::steady_clock::time_point t_entry, uint64_t launches0)
{
    const int dev = device_of(c.opt);
    ScopedDevice sd(dev);
    const double alloc0 = alloc_ms_counter();
    int rc = NNLM_OK;
    std::chrono::steady_clock::time_point t_done;
    {
    Engine eng(c.n, c.m, c.K, c.method, c.precision, dev);
    eng.timer.enable(c.opt && c.opt->verbose_timing);
    Events ev;
    ev.record(0, eng.stream());
    eng.upload_A(c.A);                                      // + missing detection and KL constant, :64-73
    eng.set_factors(c.W, c.H);                              // :82-98 (explicit init; the shim draws the default one)
    eng.set_masks(c.Wm, c.Hm);
    eng.set_penalties(c.alpha, c.beta);
    eng.set_inner(c.inner_max_iter, c.inner_rel_tol);
    ev.record(1, eng.stream());
    eng.sync();
    rc = anls_loop(eng, c, true, nullptr, t_entry, launches0);
    if (c.stats && rc == NNLM_OK) { c.stats->upload_ms = elapsed(ev.e[0], ev.e[1]); c.stats->n_gpus_used = 1; }
    t_done = std::chrono::steady_clock::now();
    }   // the engine and its device buffers go here
    if (std::getenv("NNLM_B200_TRACE"))
        std::fprintf(stderr, "[nnlm_b200] teardown: cudaFreeAsync %.1f ms, pinned alloc/free %.1f ms (thread totals)\n", free_ms_counter(), pinned_ms_counter());
    if (c.stats && rc == NNLM_OK) {
        c.stats->host_alloc_ms = alloc_ms_counter() - alloc0;
        c.stats->host_teardown_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_done).count();
    }
    return rc;
}

// N GPUs behind the one call (SURVEY.md §8b "Threading"): the calling thread drives rank 0, N-1 worker threads the other
// devices; every rank uploads its column shard (contiguous in the column-major host matrix) and its row shard (a strided
// 2-D copy) straight from the caller's A, then runs the same loop over NCCL. Worker threads never touch the caller's
// runtime (R): only rank 0 calls the interrupt and print callbacks.
int nnmf_multi(const NnmfCall& c, int R, std::chrono::steady_clock::time_point t_entry, uint64_t launches0, std::string* errmsg)
{
    const int dev0 = device_of(c.opt);
    NcclId id;
    Comm::unique_id(&id);
    Team team(R);
    std::vector<int> rcs(R, NNLM_OK);
    std::vector<std::string> msgs(R);
    auto body = [&](int rank) {
        int rc = NNLM_OK;
        bool setup_ok = false;
        try {
            const int dev = dev0 + rank;
            ScopedDevice sd(dev);
            std::unique_ptr<Comm> comm;
            std::unique_ptr<Engine> eng;
            std::unique_ptr<Events> ev;
            try {
                comm.reset(new Comm(id, rank, R, dev));                     // collective: every rank of the team calls it
                eng.reset(new Engine(c.n, c.m, c.K, c.method, c.precision, dev, true, comm.get()));
                eng->timer.enable(c.opt && c.opt->verbose_timing);
                ev.reset(new Events);
                ev->record(0, eng->stream());
                const int64_t n = c.n, m = c.m, c0 = eng->col0(), mc = eng->cols_local(), r0 = eng->row0(), nr = eng->rows_local();
                DevBuf<double> dC((size_t)n * std::max<int64_t>(mc, 1)), dR((size_t)std::max<int64_t>(nr, 1) * m);
                if (mc > 0) NNLM_CUDA_CHECK(cudaMemcpyAsync(dC.p, c.A + (size_t)n * c0, (size_t)n * mc * sizeof(double), cudaMemcpyHostToDevice, eng->stream()));
                if (nr > 0) NNLM_CUDA_CHECK(cudaMemcpy2DAsync(dR.p, (size_t)nr * sizeof(double), c.A + r0, (size_t)n * sizeof(double),
                                                              (size_t)nr * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, eng->stream()));
                eng->h2d_bytes += ((size_t)n * mc + (size_t)nr * m) * sizeof(double);
                setup_ok = true;
                team.bar.wait();                                           // (a) nobody enters the first collective unless all got here
                if (team.failed.load()) throw Error(NNLM_E_CUDA, "another rank of the multi-GPU call failed during set-up");
                eng->ingest_shards(dC.p, dR.p);
                eng->sync();
                dC.release(); dR.release();
                eng->set_factors(c.W, c.H);
                eng->set_masks(c.Wm, c.Hm);
                eng->set_penalties(c.alpha, c.beta);
                eng->set_inner(c.inner_max_iter, c.inner_rel_tol);
                ev->record(1, eng->stream());
                eng->sync();
            } catch (...) {
                if (!setup_ok) { team.failed.store(1); team.bar.wait(); }
                throw;
            }
            rc = anls_loop(*eng, c, rank == 0, &team, t_entry, launches0);
            if (rank == 0 && c.stats && rc == NNLM_OK) { c.stats->upload_ms = elapsed(ev->e[0], ev->e[1]); c.stats->n_gpus_used = R; }
            eng->sync();
            ev.reset(); eng.reset(); comm.reset();
        } catch (const Error& e) { rc = e.code; msgs[rank] = e.what(); }
        catch (const std::exception& e) { rc = NNLM_E_CUDA; msgs[rank] = e.what(); }
        rcs[rank] = rc;
    };
    std::vector<std::thread> workers;
    for (int r = 1; r < R; r++) workers.emplace_back(body, r);
    body(0);
    for (auto& t : workers) t.join();
    for (int r = 0; r < R; r++)
        if (rcs[r] != NNLM_OK) { *errmsg = "rank " + std::to_string(r) + ": " + (msgs[r].empty() ? "interrupted" : msgs[r]); return rcs[r]; }
    return NNLM_OK;
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

size_t nnlm_sizeof(int which) { return which == 0 ? sizeof(nnlm_options) : which == 1 ? sizeof(nnlm_stats) : 0; }

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
    const auto t_entry = std::chrono::steady_clock::now();
    const int rc = guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(A && W && H && alpha && beta, "nnlm_nnmf: NULL argument");
        NNLM_REQUIRE(n > 0 && m > 0 && K > 0, "nnlm_nnmf: dimensions must be positive");
        NNLM_REQUIRE(method >= 1 && method <= 4, "nnlm_nnmf: method code must be 1..4");
        const uint32_t tr = trace < 1 ? 1 : trace;                                           // src/nnmf.cpp:53
        const uint32_t err_len = (uint32_t)std::ceil((double)max_iter / (double)tr) + 1;     // :54
        NNLM_REQUIRE(mse && mkl && target && avg_epoch && err_cap >= err_len,
                     "nnlm_nnmf: error vectors must hold ceil(max_iter/trace)+1 entries");
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        NnmfCall c{A, n, m, K, W, H, Wm, Hm, alpha, beta, max_iter, rel_tol, verbose, inner_max_iter, inner_rel_tol, method, tr,
                   mse, mkl, target, avg_epoch, n_err, n_iter, converged, interrupt, interrupt_user, opt, stats,
                   precision_of(opt, n, m, true)};
        int R = (opt && opt->comm) ? 1 : gpus_requested(opt);
        R = (int)std::min<int64_t>(std::min<int64_t>(R, device_count_noexcept()), std::min(n, m));
        if (R <= 1) return nnmf_single(c, t_entry, launches0);
        std::string msg;
        const int rc2 = nnmf_multi(c, R, t_entry, launches0, &msg);
        if (rc2 != NNLM_OK) set_err(err, errlen, rc2 == NNLM_E_INTERRUPT ? "nnlm_nnmf: interrupted" : msg.c_str());
        return rc2;
    });
    if (rc == NNLM_E_INTERRUPT && err && errlen && !err[0]) set_err(err, errlen, "nnlm_nnmf: interrupted");
    if (stats && rc == NNLM_OK)
        stats->host_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count();
    return rc;
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
        ScopedDevice sd(device_of(opt));
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        Engine eng(n, m, k, method, precision_of(opt, n, m, false), opt ? opt->device : -1, /*both_sides=*/false);
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
        ScopedDevice sd(device_of(opt));
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        // update(beta, x.t(), y, mask, alpha, ...)  (src/nnlm.cpp:44-47): the engine's "W" is x (n x p), its "A" is y
        Engine eng(n, q, (int)p, method, precision_of(opt, n, q, false), opt ? opt->device : -1, /*both_sides=*/false);
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
        ScopedDevice sd(device_of(opt));
        if (stats) std::memset(stats, 0, sizeof *stats);
        const uint64_t launches0 = launch_counter().load();
        Engine eng(n, m, k, NNLM_SCD_MSE, precision_of(opt, n, m, true), opt ? opt->device : -1, /*both_sides=*/false);
        eng.set_missing_mode(0);
        eng.upload_A(A);
        std::vector<double> H0((size_t)k * m, 0.0);
        eng.set_factors_t(Wt, H0.data());
        eng.cross_only(Q);
        fill_stats(stats, eng, launches0);
        return NNLM_OK;
    });
}

// diagnostic entry: the NA-path corrections of every column through the tensor-core mask contraction (na_gram.cu)
int nnlm_na_corrections(const double* Wt, const double* A, int32_t k, int64_t n, int64_t m, double* S, int64_t s_capacity,
                        int64_t* width, const nnlm_options* opt, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(Wt && A && S && width && k > 0 && n > 0 && m > 0, "nnlm_na_corrections: bad argument");
        ScopedDevice sd(device_of(opt));
        Engine eng(n, m, k, NNLM_SCD_MSE, NNLM_PREC_FAST, opt ? opt->device : -1, /*both_sides=*/false);
        eng.set_missing_mode(1);
        eng.upload_A(A);
        std::vector<double> H0((size_t)k * m, 0.0);
        eng.set_factors_t(Wt, H0.data());
        NNLM_REQUIRE(s_capacity >= m * ((int64_t)(k * (k + 1) / 2 + k + 127) / 128 * 128), "nnlm_na_corrections: S is too small");
        *width = eng.na_corrections(S);
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
        ScopedDevice sd(device_of(opt));
        std::unique_ptr<nnlm_session> s(new nnlm_session);
        s->launches0 = launch_counter().load();
        s->eng.reset(new Engine(n, m, K, method, precision_of(opt, n, m, true), opt ? opt->device : -1));
        s->eng->timer.enable(opt && opt->verbose_timing);
        Events ev;
        ev.record(0, s->eng->stream());
        s->eng->upload_A(A);
        ev.record(1, s->eng->stream());
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
        ScopedDevice sd(device_of(opt));
        std::unique_ptr<nnlm_session> s(new nnlm_session);
        s->launches0 = launch_counter().load();
        Comm* comm = (opt && opt->comm) ? nnlm_comm_get(static_cast<nnlm_comm*>(opt->comm)) : nullptr;
        s->eng.reset(new Engine(n, m, K, method, precision_of(opt, n, m, true), opt ? opt->device : -1, true, comm));
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
        ScopedDevice sd(device_of(opt));
        std::unique_ptr<nnlm_session> s(new nnlm_session);
        s->launches0 = launch_counter().load();
        Comm* comm = (opt && opt->comm) ? nnlm_comm_get(static_cast<nnlm_comm*>(opt->comm)) : nullptr;
        s->eng.reset(new Engine(n, m, K, method, precision_of(opt, n, m, true), opt ? opt->device : -1, true, comm));
        s->eng->timer.enable(opt && opt->verbose_timing);
        Engine& e = *s->eng;
        Events ev;
        ev.record(0, e.stream());
        const size_t cc = (size_t)n * e.cols_local(), cr = (size_t)e.rows_local() * m;
        DevBuf<double> dC(std::max<size_t>(cc, 1)), dR(std::max<size_t>(cr, 1));
        NNLM_CUDA_CHECK(cudaMemcpyAsync(dC.p, Acol, cc * sizeof(double), cudaMemcpyHostToDevice, e.stream()));
        NNLM_CUDA_CHECK(cudaMemcpyAsync(dR.p, Arow, cr * sizeof(double), cudaMemcpyHostToDevice, e.stream()));
        e.h2d_bytes += (cc + cr) * sizeof(double);
        e.ingest_shards(dC.p, dR.p);
        ev.record(1, e.stream());
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
        ScopedDevice sd(s->eng->device());
        Events ev;
        cudaStream_t st = s->eng->stream();
        s->eng->sync();
        ev.record(0, st);
        for (uint32_t i = 0; i < iters; i++) { s->eng->half_w(); s->eng->half_h(); }          // src/nnmf.cpp:109-133
        ev.record(1, st);
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

int nnlm_session_mse(nnlm_session* s, double* mse, int32_t* from_identity, char* err, size_t errlen)
{
    return guarded(err, errlen, [&]() -> int {
        NNLM_REQUIRE(s && mse, "nnlm_session_mse: NULL argument");
        Engine& e = *s->eng;
        ErrorTerms t;
        e.errors(&t, /*want_kl=*/false);
        *mse = t.sum_sq / (double)(e.n() * e.m() - e.n_missing());
        if (from_identity) *from_identity = e.last_mse_from_identity() ? 1 : 0;
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
