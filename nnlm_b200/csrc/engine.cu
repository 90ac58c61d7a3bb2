// engine.cu — see engine.cuh. Host-side orchestration only: every arithmetic step is a kernel of this library.
#include "engine.cuh"

#include <algorithm>

namespace nnlm {

namespace {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev >= 0 && dev != prev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

KernelTimer::~KernelTimer()
{
    for (auto& sp : spans_) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : pool_) cudaEventDestroy(e);
}

cudaEvent_t KernelTimer::get()
{
    if (!pool_.empty()) { cudaEvent_t e = pool_.back(); pool_.pop_back(); return e; }
    cudaEvent_t e;
    NNLM_CUDA_CHECK(cudaEventCreate(&e));
    return e;
}

void KernelTimer::begin(Cat c, cudaStream_t st)
{
    if (!on_) return;
    cur_ = c;
    cur_a_ = get();
    NNLM_CUDA_CHECK(cudaEventRecord(cur_a_, st));
}

void KernelTimer::end(cudaStream_t st)
{
    if (!on_) return;
    cudaEvent_t b = get();
    NNLM_CUDA_CHECK(cudaEventRecord(b, st));
    spans_.push_back(Span{cur_a_, b, cur_});
}

void KernelTimer::collect()
{
    for (auto& sp : spans_) {
        float t = 0;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { ms[sp.c] += t; count[sp.c]++; }
        pool_.push_back(sp.a);
        pool_.push_back(sp.b);
    }
    spans_.clear();
}

Engine::Engine(int64_t n, int64_t m, int k, int method, int precision, int device, bool both_sides)
    : n_(n), m_(m), k_(k), method_(method), device_(device), both_sides_(both_sides)
{
    NNLM_REQUIRE(n > 0 && m > 0, "A must have positive dimensions");
    NNLM_REQUIRE(k >= 1, "rank k must be positive");
    NNLM_REQUIRE(method >= 1 && method <= 4, "method code must be 1..4 (R/misc.R:28-35)");
    if (method <= 2) NNLM_REQUIRE(k <= 128, "square-loss solvers support rank k <= 128");
    if (device_ < 0) NNLM_CUDA_CHECK(cudaGetDevice(&device_));
    NNLM_CUDA_CHECK(cudaSetDevice(device_));
    // precision policy (include/nnlm_b200.h): the resident copies of A are fp64 unless the fast path is requested
    precision_req_ = precision;
    storage_ = (precision == NNLM_PREC_FAST) ? Storage::F32 : Storage::F64;       // refined in ingest_device_A
    NNLM_CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    Wt_.alloc((size_t)k_ * n_);
    H_.alloc((size_t)k_ * m_);
    ensure_scratch();
}

Engine::~Engine()
{
    if (st_) { cudaStreamSynchronize(st_); cudaStreamDestroy(st_); }
}

void Engine::ensure_scratch()
{
    const int64_t big = std::max(n_, m_);
    gram_part_.alloc((size_t)gram_splits(big) * k_ * k_);
    G_.alloc((size_t)k_ * k_);
    Graw_.alloc((size_t)k_ * k_);
    sumY_.alloc(k_);
    if (method_ <= 2) {
        size_t q = (size_t)cross_simt_splits(k_, n_, m_) * k_ * m_;
        if (both_sides_) q = std::max(q, (size_t)cross_simt_splits(k_, m_, n_) * k_ * n_);
        if (cross_tc_supported(k_)) {
            plan_h_ = cross_tc_plan(k_, n_, m_);
            q = std::max(q, (size_t)plan_h_.slots * k_ * m_);
            if (both_sides_) {
                plan_w_ = cross_tc_plan(k_, m_, n_);
                q = std::max(q, (size_t)plan_w_.slots * k_ * n_);
            }
        }
        Qp_.alloc(q);
    } else {
        Yr_.alloc((size_t)k_ * big);
        size_t w = solve_kl_scratch_doubles(n_, m_);
        if (both_sides_) w = std::max(w, solve_kl_scratch_doubles(m_, n_));
        if (w) wh_.alloc(w);
    }
    size_t rp = std::max<size_t>((size_t)ingest_part_count(n_, m_) * 2, (size_t)error_part_count(n_, m_) * 2);
    rp = std::max<size_t>(rp, (size_t)stats_part_count(big) * 3);
    red_part_.alloc(rp);
    small_.alloc(16);
    sweeps_.alloc(1);
    host_small_.alloc(16);
    NNLM_CUDA_CHECK(cudaMemsetAsync(sweeps_.p, 0, sizeof(unsigned long long), st_));
}

void Engine::sync() { NNLM_CUDA_CHECK(cudaStreamSynchronize(st_)); timer.collect(); }

void Engine::upload_A(const double* A)
{
    DeviceGuard g(device_);
    const size_t cnt = (size_t)n_ * m_;
    A64_.alloc(cnt);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(A64_.p, A, cnt * sizeof(double), cudaMemcpyHostToDevice, st_));
    h2d_bytes += cnt * sizeof(double);
    ingest_device_A(nullptr);
}

void Engine::ingest_device_A(const double* dA)
{
    DeviceGuard g(device_);
    const size_t cnt = (size_t)n_ * m_;
    if (dA) {
        A64_.alloc(cnt);
        NNLM_CUDA_CHECK(cudaMemcpyAsync(A64_.p, dA, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st_));
    }
    const int64_t parts = ingest_part_count(n_, m_);
    // pass 1: missing-entry count and the constant part of the KL distance (src/nnmf.cpp:64-73)
    launch_ingest<double>(A64_.p, n_, m_, 0, m_, nullptr, nullptr, red_part_.p, st_);
    launch_reduce_partials(red_part_.p, parts, 2, small_.p, st_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(host_small_.p, small_.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += 2 * sizeof(double);
    kl_const_sum_ = host_small_.p[0];
    n_missing_ = (int64_t)host_small_.p[1];
    // pass 2: the resident copies, in the storage the precision policy selects (include/nnlm_b200.h)
    if (precision_req_ == NNLM_PREC_FAST && method_ <= 2 && cross_tc_supported(k_) && !use_missing_path())
        storage_ = Storage::F16X2;
    if (storage_ == Storage::F64) {
        if (both_sides_) {
            At64_.alloc(cnt);
            launch_ingest<double>(A64_.p, n_, m_, 0, m_, nullptr, At64_.p, red_part_.p, st_);
        }
    } else if (storage_ == Storage::F32) {
        A32_.alloc(cnt);
        if (both_sides_) At32_.alloc(cnt);
        launch_ingest<float>(A64_.p, n_, m_, 0, m_, A32_.p, both_sides_ ? At32_.p : nullptr, red_part_.p, st_);
    } else {
        const int64_t ld_n = cross_tc_ld(n_), ld_m = cross_tc_ld(m_);
        const int np = cross_tc_np(k_);
        a_hi_.alloc((size_t)ld_n * m_); a_lo_.alloc((size_t)ld_n * m_);
        if (both_sides_) { t_hi_.alloc((size_t)ld_m * n_); t_lo_.alloc((size_t)ld_m * n_); }
        const int64_t ld_f = std::max(ld_n, both_sides_ ? ld_m : (int64_t)0);
        f_hi_.alloc((size_t)np * ld_f); f_lo_.alloc((size_t)np * ld_f);
        scale_a_.alloc(1); fscales_.alloc(np); unscale_.alloc(np); rowmax_.alloc(np + 1);
        // pitch padding of the planes must read as zero (it is never written by the conversion kernels)
        NNLM_CUDA_CHECK(cudaMemsetAsync(a_hi_.p, 0, a_hi_.bytes(), st_));
        NNLM_CUDA_CHECK(cudaMemsetAsync(a_lo_.p, 0, a_lo_.bytes(), st_));
        if (both_sides_) {
            NNLM_CUDA_CHECK(cudaMemsetAsync(t_hi_.p, 0, t_hi_.bytes(), st_));
            NNLM_CUDA_CHECK(cudaMemsetAsync(t_lo_.p, 0, t_lo_.bytes(), st_));
        }
        colmean_.alloc(m_);
        if (both_sides_) rowmean_.alloc(n_);
        launch_means(A64_.p, n_, m_, colmean_.p, both_sides_ ? rowmean_.p : nullptr, st_);
        launch_absmax_scale(A64_.p, (int64_t)cnt, rowmax_.p + np, scale_a_.p, st_);
        launch_split_matrix(A64_.p, n_, m_, scale_a_.p, colmean_.p, both_sides_ ? rowmean_.p : nullptr, a_hi_.p, a_lo_.p, ld_n,
                            both_sides_ ? t_hi_.p : nullptr, both_sides_ ? t_lo_.p : nullptr, ld_m, st_);
        // the error evaluation (trace iterations only) reads an fp32 copy of A
        A32_.alloc(cnt);
        launch_ingest<float>(A64_.p, n_, m_, 0, m_, A32_.p, nullptr, red_part_.p, st_);
    }
    sync();
    if (storage_ != Storage::F64) A64_.release();
}

void Engine::set_factors(const double* W, const double* H)
{
    DeviceGuard g(device_);
    DevBuf<double> tmp((size_t)n_ * k_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(tmp.p, W, tmp.bytes(), cudaMemcpyHostToDevice, st_));
    launch_transpose_d(tmp.p, n_, k_, Wt_.p, st_);                          // inplace_trans(W), src/nnmf.cpp:90
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H_.p, H, H_.bytes(), cudaMemcpyHostToDevice, st_));
    sync();
    h2d_bytes += tmp.bytes() + H_.bytes();
}

void Engine::set_factors_t(const double* Wt, const double* H)
{
    DeviceGuard g(device_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(Wt_.p, Wt, Wt_.bytes(), cudaMemcpyHostToDevice, st_));
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H_.p, H, H_.bytes(), cudaMemcpyHostToDevice, st_));
    sync();
    h2d_bytes += Wt_.bytes() + H_.bytes();
}

void Engine::get_factors(double* W, double* H)
{
    DeviceGuard g(device_);
    DevBuf<double> tmp((size_t)n_ * k_);
    launch_transpose_d(Wt_.p, k_, n_, tmp.p, st_);                          // W.t(), src/nnmf.cpp:212
    NNLM_CUDA_CHECK(cudaMemcpyAsync(W, tmp.p, tmp.bytes(), cudaMemcpyDeviceToHost, st_));
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H, H_.p, H_.bytes(), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += tmp.bytes() + H_.bytes();
}

void Engine::get_H(double* H)
{
    DeviceGuard g(device_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H, H_.p, H_.bytes(), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += H_.bytes();
}

void Engine::set_masks(const int32_t* Wm, const int32_t* Hm)
{
    DeviceGuard g(device_);
    has_wm_ = Wm != nullptr;
    has_hm_ = Hm != nullptr;
    if (has_wm_) {
        DevBuf<int32_t> tmp((size_t)n_ * k_);
        Wm_.alloc((size_t)n_ * k_);
        NNLM_CUDA_CHECK(cudaMemcpyAsync(tmp.p, Wm, tmp.bytes(), cudaMemcpyHostToDevice, st_));
        launch_mask_to_u8_t(tmp.p, n_, k_, Wm_.p, st_);                     // inplace_trans(Wm), src/nnmf.cpp:78
        sync();
        h2d_bytes += tmp.bytes();
    }
    if (has_hm_) {
        DevBuf<int32_t> tmp((size_t)m_ * k_);
        Hm_.alloc((size_t)m_ * k_);
        NNLM_CUDA_CHECK(cudaMemcpyAsync(tmp.p, Hm, tmp.bytes(), cudaMemcpyHostToDevice, st_));
        launch_mask_to_u8(tmp.p, (int64_t)m_ * k_, Hm_.p, st_);
        sync();
        h2d_bytes += tmp.bytes();
    }
}

void Engine::set_penalties(const double* alpha, const double* beta)
{
    for (int i = 0; i < 3; i++) { alpha_[i] = alpha ? alpha[i] : 0.0; beta_[i] = beta ? beta[i] : 0.0; }
}

template <typename TA>
void Engine::run_half_t(const Half& h)
{
    const TA* A = static_cast<const TA*>(h.A);
    const bool missing = use_missing_path();
    if (method_ <= 2) {
        const int splits = cross_simt_splits(k_, h.len, h.ncol);
        timer.begin(KernelTimer::CROSS, st_);
        launch_cross_simt<TA>(h.Y, A, k_, h.len, h.ncol, splits, Qp_.p, st_);
        timer.end(st_);
        if (!missing) {
            solve_dense_ls(h, splits);
        } else {
            timer.begin(KernelTimer::GRAM, st_);
            launch_gram(h.Y, k_, h.len, nullptr, gram_part_.p, Graw_.p, st_);
            timer.end(st_);
            timer.begin(KernelTimer::SOLVE, st_);
            launch_solve_ls_missing<TA>(method_, h.X, h.Y, A, Graw_.p, Qp_.p, splits, h.mask, k_, h.len, h.ncol, h.pen,
                                        inner_max_iter_, inner_rel_tol_, sweeps_.p, st_);
            timer.end(st_);
        }
    } else {
        timer.begin(KernelTimer::GRAM, st_);
        launch_rowsum(h.Y, k_, h.len, gram_part_.p, sumY_.p, st_);                         // :27
        launch_transpose_d(h.Y, k_, h.len, Yr_.p, st_);
        timer.end(st_);
        timer.begin(KernelTimer::SOLVE, st_);
        launch_solve_kl<TA>(method_, h.X, Yr_.p, A, sumY_.p, h.mask, k_, h.len, h.ncol, h.pen, inner_max_iter_,
                            inner_rel_tol_, missing ? 1 : 0, wh_.p, sweeps_.p, st_);
        timer.end(st_);
    }
}

void Engine::solve_dense_ls(const Half& h, int splits)
{
    timer.begin(KernelTimer::GRAM, st_);
    launch_gram(h.Y, k_, h.len, h.pen, gram_part_.p, G_.p, st_);                          // update_with_missing.cpp:19-24
    timer.end(st_);
    timer.begin(KernelTimer::SOLVE, st_);
    if (method_ == 1 && scd_tpc_supported(k_))
        launch_scd_tpc(h.X, G_.p, Qp_.p, splits, h.mask, k_, h.ncol, h.pen[2], inner_max_iter_, inner_rel_tol_, sweeps_.p, st_);
    else
        launch_solve_ls(method_, h.X, G_.p, Qp_.p, splits, h.mask, k_, h.ncol, h.pen[2], inner_max_iter_, inner_rel_tol_,
                        sweeps_.p, st_);
    timer.end(st_);
}

void Engine::run_half_tc(const Half& h, bool w_side)
{
    const CrossPlan& plan = w_side ? plan_w_ : plan_h_;
    timer.begin(KernelTimer::GRAM, st_);
    launch_split_factor(h.Y, k_, h.len, plan.ld_f, plan.np, scale_a_.p, rowmax_.p, fscales_.p, unscale_.p, f_hi_.p, f_lo_.p, st_);
    launch_rowsum(h.Y, k_, h.len, gram_part_.p, sumY_.p, st_);
    timer.end(st_);
    timer.begin(KernelTimer::CROSS, st_);
    launch_cross_tc(plan, w_side ? t_hi_.p : a_hi_.p, w_side ? t_lo_.p : a_lo_.p, f_hi_.p, f_lo_.p, unscale_.p,
                    w_side ? rowmean_.p : colmean_.p, sumY_.p, Qp_.p, st_);
    timer.end(st_);
    solve_dense_ls(h, plan.slots);
}

void Engine::run_half(const Half& h)
{
    DeviceGuard g(device_);
    if (storage_ == Storage::F64) run_half_t<double>(h);
    else run_half_t<float>(h);
}

void Engine::half_w()
{
    NNLM_REQUIRE(both_sides_, "this engine was created for the H-half only");
    const void* At = storage_ == Storage::F64 ? (const void*)At64_.p : (const void*)At32_.p;
    const Half h{Wt_.p, n_, H_.p, m_, At, has_wm_ ? Wm_.p : nullptr, alpha_};
    if (storage_ == Storage::F16X2) { DeviceGuard g(device_); run_half_tc(h, true); }
    else run_half(h);
}

void Engine::half_h()
{
    const void* A = storage_ == Storage::F64 ? (const void*)A64_.p : (const void*)A32_.p;
    const Half h{H_.p, m_, Wt_.p, n_, A, has_hm_ ? Hm_.p : nullptr, beta_};
    if (storage_ == Storage::F16X2) { DeviceGuard g(device_); run_half_tc(h, false); }
    else run_half(h);
}

void Engine::cross_only(double* Q_host)
{
    DeviceGuard g(device_);
    int splits;
    if (storage_ == Storage::F16X2) {
        launch_split_factor(Wt_.p, k_, n_, plan_h_.ld_f, plan_h_.np, scale_a_.p, rowmax_.p, fscales_.p, unscale_.p, f_hi_.p, f_lo_.p, st_);
        launch_rowsum(Wt_.p, k_, n_, gram_part_.p, sumY_.p, st_);
        launch_cross_tc(plan_h_, a_hi_.p, a_lo_.p, f_hi_.p, f_lo_.p, unscale_.p, colmean_.p, sumY_.p, Qp_.p, st_);
        splits = plan_h_.slots;
    } else {
        splits = cross_simt_splits(k_, n_, m_);
        if (storage_ == Storage::F64) launch_cross_simt<double>(Wt_.p, A64_.p, k_, n_, m_, splits, Qp_.p, st_);
        else launch_cross_simt<float>(Wt_.p, A32_.p, k_, n_, m_, splits, Qp_.p, st_);
    }
    const size_t per = (size_t)k_ * m_;
    std::vector<double> tmp(per * splits);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), Qp_.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += tmp.size() * sizeof(double);
    for (size_t e = 0; e < per; e++) {
        double s = 0;
        for (int sp = 0; sp < splits; sp++) s += tmp[(size_t)sp * per + e];
        Q_host[e] = s;
    }
}

void Engine::errors(ErrorTerms* out)
{
    DeviceGuard g(device_);
    if (storage_ == Storage::F64) launch_error<double>(A64_.p, Wt_.p, H_.p, k_, n_, m_, red_part_.p, small_.p, st_);
    else launch_error<float>(A32_.p, Wt_.p, H_.p, k_, n_, m_, red_part_.p, small_.p, st_);
    launch_factor_stats(Wt_.p, k_, n_, red_part_.p, small_.p + 2, st_);
    launch_factor_stats(H_.p, k_, m_, red_part_.p, small_.p + 5, st_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(host_small_.p, small_.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += 8 * sizeof(double);
    out->sum_sq = host_small_.p[0];
    out->sum_kl = host_small_.p[1];
    for (int i = 0; i < 3; i++) { out->w_stats[i] = host_small_.p[2 + i]; out->h_stats[i] = host_small_.p[5 + i]; }
}

uint64_t Engine::take_sweeps()
{
    DeviceGuard g(device_);
    unsigned long long* hp = reinterpret_cast<unsigned long long*>(host_small_.p + 12);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(hp, sweeps_.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st_));
    NNLM_CUDA_CHECK(cudaMemsetAsync(sweeps_.p, 0, sizeof(unsigned long long), st_));
    sync();
    d2h_bytes += sizeof(unsigned long long);
    return *hp;
}

}  // namespace nnlm
