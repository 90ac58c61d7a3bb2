// engine.cu — see engine.cuh. Host-side orchestration only: every arithmetic step is a kernel of this library.
#include "engine.cuh"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace nnlm {

namespace {
constexpr int SCD_WARP_MAX_COLS = 0;      // columns per launch up to which the warp-per-column SCD solver is used (measured: see solve_dense_ls)
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

Engine::Engine(int64_t n, int64_t m, int k, int method, int precision, int device, bool both_sides, Comm* comm)
    : n_(n), m_(m), k_(k), method_(method), device_(device), both_sides_(both_sides), comm_(comm)
{
    NNLM_REQUIRE(n > 0 && m > 0, "A must have positive dimensions");
    NNLM_REQUIRE(k >= 1, "rank k must be positive");
    NNLM_REQUIRE(method >= 1 && method <= 4, "method code must be 1..4 (R/misc.R:28-35)");
    // the real per-method limits, checked once with a clear message (every kernel below is instantiated up to them)
    if (method <= 2) NNLM_REQUIRE(k <= 128, "nnlm_b200: the square-loss solvers (methods 'scd'/'lee' with loss 'mse') support rank / "
                                            "predictor count k <= 128");
    else NNLM_REQUIRE(k <= 256, "nnlm_b200: the KL solvers (loss 'mkl') support rank k <= 256");
    if (device_ < 0) NNLM_CUDA_CHECK(cudaGetDevice(&device_));
    DeviceGuard guard(device_);              // the caller's current device is restored on return
    const int R = comm_ ? comm_->nranks() : 1, rank = comm_ ? comm_->rank() : 0;
    NNLM_REQUIRE(n >= R && m >= R, "fewer rows or columns than ranks");
    chunk_n_ = ceil_div(n_, R); chunk_m_ = ceil_div(m_, R);
    r0_ = std::min<int64_t>(n_, rank * chunk_n_); nr_ = std::min<int64_t>(n_, r0_ + chunk_n_) - r0_;
    c0_ = std::min<int64_t>(m_, rank * chunk_m_); mc_ = std::min<int64_t>(m_, c0_ + chunk_m_) - c0_;
    // precision policy (include/nnlm_b200.h): the resident copies of A are fp64 unless the fast path is requested
    precision_req_ = precision;
    storage_ = (precision == NNLM_PREC_FAST) ? Storage::F32 : Storage::F64;       // refined in ingest_shards
    NNLM_CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    NNLM_CUDA_CHECK(cudaStreamCreateWithFlags(&st2_, cudaStreamNonBlocking));
    NNLM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    NNLM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    Wt_.alloc((size_t)k_ * chunk_n_ * R);
    H_.alloc((size_t)k_ * chunk_m_ * R);
    NNLM_CUDA_CHECK(cudaMemsetAsync(Wt_.p, 0, Wt_.bytes(), st_));
    NNLM_CUDA_CHECK(cudaMemsetAsync(H_.p, 0, H_.bytes(), st_));
    ensure_scratch();
}

Engine::~Engine()
{
    const bool trace = std::getenv("NNLM_B200_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    const auto t0 = now();
    if (st2_) { cudaStreamSynchronize(st2_); cudaStreamDestroy(st2_); }
    if (st_) { cudaStreamSynchronize(st_); cudaStreamDestroy(st_); }
    const auto t1 = now();
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
    if (trace)
        std::fprintf(stderr, "[nnlm_b200] ~Engine: streams %.1f ms, events %.1f ms (buffers are released after this)\n",
                     std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(now() - t1).count());
}

void Engine::ensure_scratch()
{
    const int64_t big = std::max(n_, m_);
    gram_part_.alloc((size_t)gram_splits(big) * k_ * k_);
    rowsum_part_.alloc((size_t)gram_splits(big) * k_ * 2);
    G_.alloc((size_t)k_ * k_);
    Graw_.alloc((size_t)k_ * k_);
    G2_.alloc((size_t)k_ * k_);
    sumY_.alloc(k_);
    if (method_ <= 2) {
        size_t q = (size_t)cross_simt_splits(k_, n_, mc_) * k_ * std::max<int64_t>(mc_, 1);
        if (both_sides_) q = std::max(q, (size_t)cross_simt_splits(k_, m_, nr_) * k_ * std::max<int64_t>(nr_, 1));
        if (cross_tc_supported(k_)) {
            if (mc_ > 0) { plan_h_ = cross_tc_plan(k_, n_, mc_); q = std::max(q, (size_t)plan_h_.slots * k_ * mc_); }
            if (both_sides_ && nr_ > 0) { plan_w_ = cross_tc_plan(k_, m_, nr_); q = std::max(q, (size_t)plan_w_.slots * k_ * nr_); }
        }
        Qp_.alloc(q);
    } else {
        Yr_.alloc((size_t)k_ * big);
        Yr32_.alloc((size_t)k_ * big);
        size_t w = solve_kl_scratch_doubles(n_, mc_);
        if (both_sides_) w = std::max(w, solve_kl_scratch_doubles(m_, nr_));
        if (w) wh_.alloc(w);
    }
    size_t rp = std::max<size_t>((size_t)ingest_part_count(n_, std::max<int64_t>(mc_, 1)) * INGEST_PART_WIDTH,
                                 (size_t)error_part_count(n_, std::max<int64_t>(mc_, 1)) * 2);
    if (both_sides_) rp = std::max<size_t>(rp, (size_t)ingest_part_count(std::max<int64_t>(nr_, 1), m_) * INGEST_PART_WIDTH);
    rp = std::max<size_t>(rp, (size_t)stats_part_count(big) * 3);
    red_part_.alloc(rp);
    small_.alloc(16);
    tpc_scratch_.alloc(scd_tpc_scratch_doubles());      // [0]: the solver's group counter, [1]: the ticket of k_factor_prep
    NNLM_CUDA_CHECK(cudaMemsetAsync(tpc_scratch_.p, 0, tpc_scratch_.bytes(), st_));
    sweeps_.alloc(1);
    NNLM_CUDA_CHECK(cudaMemsetAsync(sweeps_.p, 0, sizeof(unsigned long long), st_));
}

void Engine::sync() { NNLM_CUDA_CHECK(cudaStreamSynchronize(st_)); timer.collect(); }

void Engine::upload_A(const double* A)
{
    DeviceGuard g(device_);
    NNLM_REQUIRE(comm_ == nullptr, "upload_A takes the whole matrix: use ingest_shards on the sharded path");
    const size_t cnt = (size_t)n_ * m_;
    DevBuf<double> dA(cnt);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, cnt * sizeof(double), cudaMemcpyHostToDevice, st_));
    h2d_bytes += cnt * sizeof(double);
    if (precision_req_ != NNLM_PREC_FAST) {          // fp64 storage keeps the uploaded buffer as the column copy
        A64_ = std::move(dA);
        ingest_shards(nullptr, both_sides_ ? A64_.p : nullptr);
    } else {
        ingest_shards(dA.p, both_sides_ ? dA.p : nullptr);
    }
}

// dAcol == nullptr means "A64_ already holds the column copy"
void Engine::ingest_shards(const double* dAcol, const double* dArow)
{
    DeviceGuard g(device_);
    const size_t cnt_c = (size_t)n_ * mc_, cnt_r = (size_t)nr_ * m_;
    if (!dAcol) dAcol = A64_.p;
    NNLM_REQUIRE(!both_sides_ || dArow != nullptr, "the row shard of A is required when both halves run");
    // pass 1: missing-entry count and the constant part of the KL distance (src/nnmf.cpp:64-73), over the column shards
    double* acc = small_.p;
    NNLM_CUDA_CHECK(cudaMemsetAsync(acc, 0, INGEST_PART_WIDTH * sizeof(double), st_));
    if (mc_ > 0) {
        launch_ingest<double>(dAcol, n_, mc_, 0, mc_, nullptr, nullptr, red_part_.p, st_);
        launch_reduce_partials(red_part_.p, ingest_part_count(n_, mc_), INGEST_PART_WIDTH, acc, st_);
    }
    if (comm_) comm_->allreduce_sum_f64(acc, INGEST_PART_WIDTH, st_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(host_small_.p, acc, INGEST_PART_WIDTH * sizeof(double), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += INGEST_PART_WIDTH * sizeof(double);
    kl_const_sum_ = host_small_.p[0];
    n_missing_ = (int64_t)host_small_.p[1];
    sum_sq_a_ = host_small_.p[2];
    q_valid_ = false;
    // pass 2: the resident copies, in the storage the precision policy selects (include/nnlm_b200.h)
    if (precision_req_ == NNLM_PREC_AUTO) {
        // small problems gain nothing from reduced storage; large dense square-loss problems take the tensor-core planes
        // unless the data is so heavy-tailed that single entries rival whole column sums: the fp32 TMEM accumulator then
        // loses the small products added after a spike (measured 5e-6 of sum|F||A| with one entry at 1e4 x rms,
        // tests/test_gpu_scale_parity.py). Entries up to rms * min(n, m) / 16 keep that loss below ~1e-7.
        const bool large = (double)n_ * (double)m_ >= 4.0e6;
        int choice = NNLM_PREC_EXACT;
        if (large) {
            choice = NNLM_PREC_FAST;
            if (method_ <= 2 && cross_tc_supported(k_) && !use_missing_path()) {
                unsigned long long* mx = reinterpret_cast<unsigned long long*>(small_.p + 8);
                launch_absmax(dAcol, (int64_t)cnt_c, mx, st_);
                if (comm_) comm_->allreduce_max_u64(mx, 1, st_);
                NNLM_CUDA_CHECK(cudaMemcpyAsync(host_small_.p + 8, mx, sizeof(double), cudaMemcpyDeviceToHost, st_));
                sync();
                d2h_bytes += sizeof(double);
                const double amax = host_small_.p[8];          // the bit pattern of max |A| IS that double
                const double finite = (double)n_ * (double)m_ - (double)n_missing_;
                const double rms = finite > 0 ? std::sqrt(sum_sq_a_ / finite) : 0.0;
                if (amax > rms * (double)std::min(n_, m_) / 16.0) choice = NNLM_PREC_EXACT;
            }
        }
        precision_req_ = choice;
        storage_ = choice == NNLM_PREC_FAST ? Storage::F32 : Storage::F64;
    }
    // fast square-loss path: fp16 planes for the tensor-core cross-product; with missing entries the planes hold zero there
    // (= the column mean after centring) and the NA corrections come from the mask contraction of na_gram.cu
    if (precision_req_ == NNLM_PREC_FAST && method_ <= 2 && cross_tc_supported(k_)
        && (!use_missing_path() || std::getenv("NNLM_NA_FP64") == nullptr))
        storage_ = Storage::F16X2;
    if (storage_ == Storage::F64) {
        if (dAcol != A64_.p) {
            A64_.alloc(cnt_c);
            NNLM_CUDA_CHECK(cudaMemcpyAsync(A64_.p, dAcol, cnt_c * sizeof(double), cudaMemcpyDeviceToDevice, st_));
        }
        if (both_sides_ && nr_ > 0) {
            At64_.alloc(cnt_r);
            launch_ingest<double>(dArow, nr_, m_, 0, m_, nullptr, At64_.p, red_part_.p, st_);
        }
    } else if (storage_ == Storage::F32) {
        A32_.alloc(cnt_c);
        if (mc_ > 0) launch_ingest<float>(dAcol, n_, mc_, 0, mc_, A32_.p, nullptr, red_part_.p, st_);
        if (both_sides_ && nr_ > 0) {
            At32_.alloc(cnt_r);
            launch_ingest<float>(dArow, nr_, m_, 0, m_, nullptr, At32_.p, red_part_.p, st_);
        }
    } else {
        const int64_t ld_n = cross_tc_ld(n_), ld_m = cross_tc_ld(m_);
        const int np = cross_tc_np(k_);
        scale_a_.alloc(1); fscales_.alloc(np); unscale_.alloc(np); rowmax_.alloc(np + 1);
        const int64_t ld_f = std::max(ld_n, both_sides_ ? ld_m : (int64_t)0);
        f_hi_.alloc((size_t)np * ld_f); f_lo_.alloc((size_t)np * ld_f);
        // one power-of-two scale for the whole matrix: max |A| over all shards
        launch_absmax(dAcol, (int64_t)cnt_c, rowmax_.p + np, st_);
        if (comm_) comm_->allreduce_max_u64(rowmax_.p + np, 1, st_);
        launch_scale_from_max(rowmax_.p + np, scale_a_.p, st_);
        if (mc_ > 0) {
            a_hi_.alloc((size_t)ld_n * mc_); a_lo_.alloc((size_t)ld_n * mc_); colmean_.alloc(mc_);
            // pitch padding of the planes must read as zero (it is never written by the conversion kernels)
            NNLM_CUDA_CHECK(cudaMemsetAsync(a_hi_.p, 0, a_hi_.bytes(), st_));
            NNLM_CUDA_CHECK(cudaMemsetAsync(a_lo_.p, 0, a_lo_.bytes(), st_));
            launch_means(dAcol, n_, mc_, colmean_.p, nullptr, st_);
            launch_split_matrix(dAcol, n_, mc_, scale_a_.p, colmean_.p, nullptr, a_hi_.p, a_lo_.p, ld_n, nullptr, nullptr, 0, st_);
            // the error evaluation (trace iterations only) reads an fp32 copy of the column shard
            A32_.alloc(cnt_c);
            launch_ingest<float>(dAcol, n_, mc_, 0, mc_, A32_.p, nullptr, red_part_.p, st_);
        }
        if (both_sides_ && nr_ > 0) {
            t_hi_.alloc((size_t)ld_m * nr_); t_lo_.alloc((size_t)ld_m * nr_); rowmean_.alloc(nr_);
            NNLM_CUDA_CHECK(cudaMemsetAsync(t_hi_.p, 0, t_hi_.bytes(), st_));
            NNLM_CUDA_CHECK(cudaMemsetAsync(t_lo_.p, 0, t_lo_.bytes(), st_));
            launch_means(dArow, nr_, m_, nullptr, rowmean_.p, st_);
            launch_split_matrix(dArow, nr_, m_, scale_a_.p, nullptr, rowmean_.p, nullptr, nullptr, 0, t_hi_.p, t_lo_.p, ld_m, st_);
        }
        if (use_missing_path()) {
            // the missing indicator as fp16 0/1 planes, and the scratch of the mask contraction
            const int64_t pt = na_packed_width(k_);
            size_t q2 = 0, sN = 0;
            // the mask contraction runs on CTA pairs (k_mask_tc2: 131 flop per byte through L2 instead of 87);
            // NNLM_NA_PAIRS=0 selects the single-CTA kernel (experiments)
            static const bool na_pairs = [] { const char* e = std::getenv("NNLM_NA_PAIRS"); return !(e && atoi(e) == 0); }();
            if (mc_ > 0) {
                mk_.alloc((size_t)ld_n * mc_);
                NNLM_CUDA_CHECK(cudaMemsetAsync(mk_.p, 0, mk_.bytes(), st_));
                launch_mask_planes<double>(dAcol, n_, mc_, mk_.p, ld_n, nullptr, 0, st_);
                plan_na_h_ = cross_tc_plan(NA_TILE, n_, mc_, na_pairs);
                q2 = (size_t)plan_na_h_.slots * mc_ * NA_TILE; sN = (size_t)mc_ * pt;
            }
            if (both_sides_ && nr_ > 0) {
                mkt_.alloc((size_t)ld_m * nr_);
                NNLM_CUDA_CHECK(cudaMemsetAsync(mkt_.p, 0, mkt_.bytes(), st_));
                launch_mask_planes<double>(dArow, nr_, m_, nullptr, 0, mkt_.p, ld_m, st_);
                plan_na_w_ = cross_tc_plan(NA_TILE, m_, nr_, na_pairs);
                q2 = std::max(q2, (size_t)plan_na_w_.slots * nr_ * NA_TILE); sN = std::max(sN, (size_t)nr_ * pt);
            }
            zplanes_.alloc((size_t)NA_SLICES * NA_TILE * ld_f);
            zunscale_.alloc((size_t)NA_SLICES * NA_TILE);
            Qp2_.alloc(q2);
            S_.alloc(sN);
        }
    }
    sync();
    if (storage_ != Storage::F64) A64_.release();
}

void Engine::set_factors(const double* W, const double* H)
{
    DeviceGuard g(device_);
    DevBuf<double> tmp((size_t)n_ * k_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(tmp.p, W, tmp.bytes(), cudaMemcpyHostToDevice, st_));
    q_valid_ = false;
    launch_transpose_d(tmp.p, n_, k_, Wt_.p, st_);                          // inplace_trans(W), src/nnmf.cpp:90
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H_.p, H, sizeof(double) * k_ * m_, cudaMemcpyHostToDevice, st_));
    sync();
    h2d_bytes += tmp.bytes() + sizeof(double) * k_ * m_;
}

void Engine::set_factors_t(const double* Wt, const double* H)
{
    DeviceGuard g(device_);
    q_valid_ = false;
    NNLM_CUDA_CHECK(cudaMemcpyAsync(Wt_.p, Wt, sizeof(double) * k_ * n_, cudaMemcpyHostToDevice, st_));
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H_.p, H, sizeof(double) * k_ * m_, cudaMemcpyHostToDevice, st_));
    sync();
    h2d_bytes += sizeof(double) * k_ * (n_ + m_);
}

void Engine::get_factors(double* W, double* H)
{
    DeviceGuard g(device_);
    DevBuf<double> tmp((size_t)n_ * k_);
    launch_transpose_d(Wt_.p, k_, n_, tmp.p, st_);                          // W.t(), src/nnmf.cpp:212
    NNLM_CUDA_CHECK(cudaMemcpyAsync(W, tmp.p, tmp.bytes(), cudaMemcpyDeviceToHost, st_));
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H, H_.p, sizeof(double) * k_ * m_, cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += tmp.bytes() + sizeof(double) * k_ * m_;
}

void Engine::get_H(double* H)
{
    DeviceGuard g(device_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(H, H_.p, sizeof(double) * k_ * m_, cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += sizeof(double) * k_ * m_;
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

// The Gram of the fixed factor (src/update_with_missing.cpp:19): each rank forms the Gram of the slice it solved in the
// previous half, the k x k partials are all-reduced (the one exchange BASELINE.json's north star names), then regularised.
// It runs on the side stream, forked from the main one, and is joined right before the solver: at 8 GPUs the Gram chain
// and its all-reduce were ~70 us of every half-iteration on the critical path.
void Engine::fork_gram(const Half& h, bool raw_only)
{
    NNLM_CUDA_CHECK(cudaEventRecord(ev_fork_, st_));
    NNLM_CUDA_CHECK(cudaStreamWaitEvent(st2_, ev_fork_, 0));
    timer.begin(KernelTimer::GRAM, st2_);
    launch_gram(h.Yloc, k_, h.len_loc, nullptr, gram_part_.p, Graw_.p, st2_);
    timer.end(st2_);
    if (comm_) {
        timer.begin(KernelTimer::COMM, st2_);
        comm_->allreduce_sum_f64(Graw_.p, (size_t)k_ * k_, st2_);
        comm_bytes += sizeof(double) * k_ * k_;
        timer.end(st2_);
    }
    if (!raw_only) launch_gram_regularise(Graw_.p, k_, h.pen, G_.p, st2_);                // update_with_missing.cpp:20-24
    NNLM_CUDA_CHECK(cudaEventRecord(ev_join_, st2_));
}

void Engine::join_gram() { NNLM_CUDA_CHECK(cudaStreamWaitEvent(st_, ev_join_, 0)); }

void Engine::solve_dense_ls(const Half& h, int splits)
{
    join_gram();
    timer.begin(KernelTimer::SOLVE, st_);
    if (h.ncol > 0) {
        // Few columns (small shards of the multi-GPU path): the blocked DMMA solver is bound by the latency of ONE tile
        // (350 blocks of ~1650 cycles however few tiles there are), the warp-per-column solver by 2500 dependent steps of
        // ~60 cycles per sweep set but with every column in flight at once: below the measured switch point it wins.
        static const int64_t warp_max = [] { const char* e = std::getenv("NNLM_SCD_WARP_MAX"); return e ? atoll(e) : (int64_t)SCD_WARP_MAX_COLS; }();
        if (method_ == 1 && scd_tpc_supported(k_) && h.ncol > warp_max)
            launch_scd_tpc(h.X, G_.p, Qp_.p, splits, h.mask, k_, h.ncol, h.pen[2], inner_max_iter_, inner_rel_tol_, sweeps_.p,
                           tpc_scratch_.p, st_);
        else
            launch_solve_ls(method_, h.X, G_.p, Qp_.p, splits, h.mask, k_, h.ncol, h.pen[2], inner_max_iter_, inner_rel_tol_,
                            sweeps_.p, st_);
    }
    timer.end(st_);
}

template <typename TA>
void Engine::run_half_t(const Half& h)
{
    const TA* A = static_cast<const TA*>(h.A);
    const bool missing = use_missing_path();
    if (method_ <= 2) {
        const int splits = cross_simt_splits(k_, h.len, h.ncol);
        fork_gram(h, missing);
        timer.begin(KernelTimer::CROSS, st_);
        if (h.ncol > 0) launch_cross_simt<TA>(h.Y, A, k_, h.len, h.ncol, splits, Qp_.p, st_);
        timer.end(st_);
        if (!missing) {
            solve_dense_ls(h, splits);
        } else {
            join_gram();
            timer.begin(KernelTimer::SOLVE, st_);
            launch_solve_ls_missing<TA>(method_, h.X, h.Y, A, Graw_.p, Qp_.p, splits, h.mask, k_, h.len, h.ncol, h.pen,
                                        inner_max_iter_, inner_rel_tol_, sweeps_.p, st_);
            timer.end(st_);
        }
    } else {
        // fast storage, dense A: the cluster kernel with the column state on chip (solve_kl_fast.cu); otherwise the fp64 kernel
        const bool fast_kl = std::is_same<TA, float>::value && !missing && solve_kl_fast_supported(k_, h.len)
                             && std::getenv("NNLM_KL_SLOW") == nullptr;
        timer.begin(KernelTimer::GRAM, st_);
        launch_rowsum(h.Y, k_, h.len, rowsum_part_.p, sumY_.p, st_);                       // :27
        if (fast_kl) launch_factor_rows_f32(h.Y, k_, h.len, Yr32_.p, st_);
        else launch_transpose_d(h.Y, k_, h.len, Yr_.p, st_);
        // the solvers start every column from wh = Y' x: formed for all columns at once on the tensor cores (fp16 hi/lo planes
        // of both factors, the scheme of the loss evaluation) instead of k passes over the factor rows per column group
        const float* wh0 = nullptr;
        if (fast_kl && error_tc_supported(k_) && h.ncol > 0 && std::getenv("NNLM_KL_NO_PRODUCT") == nullptr) {
            const int kp = error_tc_kp(k_);
            ew_hi_.ensure((size_t)h.len * kp); ew_lo_.ensure((size_t)h.len * kp); rsw_.ensure(4);
            eh_hi_.ensure((size_t)h.ncol * kp); eh_lo_.ensure((size_t)h.ncol * kp); rsh_.ensure(4);
            wh32_.ensure((size_t)h.len * h.ncol);
            unsigned long long* mxb = reinterpret_cast<unsigned long long*>(small_.p + 10);
            launch_split_rows(h.Y, k_, h.len, ew_hi_.p, ew_lo_.p, rsw_.p, mxb, st_);
            launch_split_rows(h.X, k_, h.ncol, eh_hi_.p, eh_lo_.p, rsh_.p, mxb + 1, st_);
            launch_product_tc(h.len, h.ncol, k_, ew_hi_.p, ew_lo_.p, rsw_.p, eh_hi_.p, eh_lo_.p, rsh_.p, wh32_.p, st_);
            wh0 = wh32_.p;
        }
        timer.end(st_);
        timer.begin(KernelTimer::SOLVE, st_);
        if (fast_kl)
            launch_solve_kl_fast(method_, h.X, Yr32_.p, reinterpret_cast<const float*>(A), wh0, sumY_.p, h.mask, k_, h.len, h.ncol, h.pen,
                                 inner_max_iter_, inner_rel_tol_, sweeps_.p, st_);
        else
            launch_solve_kl<TA>(method_, h.X, Yr_.p, A, sumY_.p, h.mask, k_, h.len, h.ncol, h.pen, inner_max_iter_,
                                inner_rel_tol_, missing ? 1 : 0, wh_.p, sweeps_.p, st_);
        timer.end(st_);
    }
}

void Engine::run_half_tc(const Half& h)
{
    const CrossPlan& plan = h.w_side ? plan_w_ : plan_h_;
    const bool missing = use_missing_path();
    fork_gram(h, missing);
    if (h.ncol > 0) {
        timer.begin(KernelTimer::GRAM, st_);
        launch_factor_prep(h.Y, k_, h.len, plan.ld_f, plan.np, scale_a_.p, rowsum_part_.p, gram_splits(h.len),
                           reinterpret_cast<unsigned int*>(tpc_scratch_.p) + 1, sumY_.p, fscales_.p, unscale_.p, f_hi_.p, f_lo_.p, st_);
        timer.end(st_);
        timer.begin(KernelTimer::CROSS, st_);
        // NA path: the per-column Grams of a near-rank-one factor amplify the cross-product's error ~750x at config 4's size
        // (T = 1 from the tiny init: rel H 1.55e-5 against the oracle with 256 indices per TMEM accumulation, scratch/na_t1_full.py),
        // and the cross-product is 3 % of that path's time: it drains every 64 indices there
        launch_cross_tc(plan, h.w_side ? t_hi_.p : a_hi_.p, h.w_side ? t_lo_.p : a_lo_.p, f_hi_.p, f_lo_.p, unscale_.p,
                        h.w_side ? rowmean_.p : colmean_.p, sumY_.p, Qp_.p, st_, missing ? 1 : 0);
        timer.end(st_);
    }
    if (!missing) { solve_dense_ls(h, plan.slots); return; }
    // NA path (src/update_with_missing.cpp:58-139): the cross-product above read missing entries as the column mean; the mask
    // contraction delivers, per column, the Gram of the missing rows and their row sums, and the solver assembles
    // G_j = G - S_j and q_j = Q_j - mean_j * (masked row sums) on the fly
    if (h.ncol > 0) {
        const CrossPlan& pna = h.w_side ? plan_na_w_ : plan_na_h_;
        timer.begin(KernelTimer::GRAM, st_);
        launch_na_gram_tc(pna, h.Y, k_, h.w_side ? mkt_.p : mk_.p, rowmax_.p, zplanes_.p, zunscale_.p, Qp2_.p, S_.p, st_);
        timer.end(st_);
    }
    join_gram();
    timer.begin(KernelTimer::SOLVE, st_);
    if (h.ncol > 0)
        launch_solve_ls_missing_packed(method_, h.X, Graw_.p, S_.p, na_packed_width(k_), Qp_.p, plan.slots,
                                       h.w_side ? rowmean_.p : colmean_.p, h.mask, k_, h.ncol, h.pen, inner_max_iter_,
                                       inner_rel_tol_, sweeps_.p, st_);
    timer.end(st_);
}

void Engine::run_half(const Half& h)
{
    DeviceGuard g(device_);
    if (storage_ == Storage::F16X2) run_half_tc(h);
    else if (storage_ == Storage::F64) run_half_t<double>(h);
    else run_half_t<float>(h);
}

// all-gather of the slices every rank just solved (in place: rank g's slice sits at g * chunk_cols columns)
void Engine::gather(double* full, int64_t chunk_cols)
{
    if (!comm_) return;
    timer.begin(KernelTimer::COMM, st_);
    const size_t cnt = (size_t)k_ * chunk_cols;
    comm_->allgather_f64(full + (size_t)comm_->rank() * cnt, full, cnt, st_);
    comm_bytes += sizeof(double) * cnt * (comm_->nranks() - 1);
    timer.end(st_);
}

void Engine::half_w()
{
    NNLM_REQUIRE(both_sides_, "this engine was created for the H-half only");
    DeviceGuard g(device_);
    q_valid_ = false;
    const void* At = storage_ == Storage::F64 ? (const void*)At64_.p : (const void*)At32_.p;
    // solve W[:, rows of this rank] given the whole H; the Gram partial is over the H columns this rank solved last
    run_half(Half{Wt_.p + (size_t)k_ * r0_, nr_, H_.p, m_, H_.p + (size_t)k_ * c0_, mc_, At,
                  has_wm_ ? Wm_.p + (size_t)k_ * r0_ : nullptr, alpha_, true});
    gather(Wt_.p, chunk_n_);
}

void Engine::half_h()
{
    DeviceGuard g(device_);
    const void* A = storage_ == Storage::F64 ? (const void*)A64_.p : (const void*)A32_.p;
    run_half(Half{H_.p + (size_t)k_ * c0_, mc_, Wt_.p, n_, Wt_.p + (size_t)k_ * r0_, nr_, A,
                  has_hm_ ? Hm_.p + (size_t)k_ * c0_ : nullptr, beta_, false});
    gather(H_.p, chunk_m_);
    // the dense square-loss H-half leaves WtA (split-K slots in Qp_) and the raw Gram of W (Graw_) behind: Engine::errors
    q_valid_ = method_ <= 2 && !use_missing_path();
    q_slots_ = storage_ == Storage::F16X2 ? plan_h_.slots : cross_simt_splits(k_, n_, mc_);
}

void Engine::cross_only(double* Q_host)
{
    DeviceGuard g(device_);
    NNLM_REQUIRE(comm_ == nullptr, "cross_only is a single-GPU diagnostic");
    int splits;
    if (storage_ == Storage::F16X2) {
        launch_split_factor(Wt_.p, k_, n_, plan_h_.ld_f, plan_h_.np, scale_a_.p, rowmax_.p, fscales_.p, unscale_.p, f_hi_.p, f_lo_.p, st_);
        launch_rowsum(Wt_.p, k_, n_, rowsum_part_.p, sumY_.p, st_);
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
    if (storage_ == Storage::F16X2 && use_missing_path()) {
        // the planes read a missing entry as its column mean: take mean_j * (row sums of the factor over the missing rows) out
        // again, exactly as the NA solver does (the masked product of src/update_with_missing.cpp:91)
        const int64_t pt = na_packed_width(k_);
        launch_na_gram_tc(plan_na_h_, Wt_.p, k_, mk_.p, rowmax_.p, zplanes_.p, zunscale_.p, Qp2_.p, S_.p, st_);
        std::vector<double> S((size_t)m_ * pt), cm(m_);
        NNLM_CUDA_CHECK(cudaMemcpyAsync(S.data(), S_.p, S.size() * sizeof(double), cudaMemcpyDeviceToHost, st_));
        NNLM_CUDA_CHECK(cudaMemcpyAsync(cm.data(), colmean_.p, cm.size() * sizeof(double), cudaMemcpyDeviceToHost, st_));
        sync();
        d2h_bytes += (S.size() + cm.size()) * sizeof(double);
        const int kk2 = k_ * (k_ + 1) / 2;
        for (int64_t j = 0; j < m_; j++)
            for (int a = 0; a < k_; a++) Q_host[a + (size_t)k_ * j] -= cm[j] * S[(size_t)j * pt + kk2 + a];
    }
}

int64_t Engine::na_corrections(double* S_host)
{
    DeviceGuard g(device_);
    NNLM_REQUIRE(comm_ == nullptr && storage_ == Storage::F16X2 && use_missing_path(),
                 "na_corrections needs the fast-precision NA path on one GPU");
    const int64_t pt = na_packed_width(k_);
    launch_na_gram_tc(plan_na_h_, Wt_.p, k_, mk_.p, rowmax_.p, zplanes_.p, zunscale_.p, Qp2_.p, S_.p, st_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(S_host, S_.p, sizeof(double) * (size_t)m_ * pt, cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += sizeof(double) * (size_t)m_ * pt;
    return pt;
}

void Engine::errors(ErrorTerms* out, bool want_kl)
{
    DeviceGuard g(device_);
    timer.begin(KernelTimer::ERROR, st_);
    NNLM_CUDA_CHECK(cudaMemsetAsync(small_.p, 0, 2 * sizeof(double), st_));
    const bool identity = !want_kl && q_valid_;
    // the fast square-loss storage evaluates both losses on the tensor cores (error_tc.cu); its KL sum includes the constant term
    const bool tc_error = storage_ == Storage::F16X2 && error_tc_supported(k_) && std::getenv("NNLM_ERR_FP64") == nullptr;
    if (identity) {
        // sum (A - W'H)^2 = ||A||^2 - 2 <H, WtA> + <WtW, HHt>: WtA is what the H-half just contracted (its split-K slots are
        // still in Qp_), WtW its raw Gram; only the k x k Gram of the new H is formed here. All fp64, O(k m + k^2).
        if (mc_ > 0) launch_dot_factor_cross(H_.p + (size_t)k_ * c0_, Qp_.p, q_slots_, k_, mc_, red_part_.p, small_.p, st_);
        if (comm_) comm_->allreduce_sum_f64(small_.p, 1, st_);
        launch_gram(H_.p, k_, m_, nullptr, gram_part_.p, G2_.p, st_);
        launch_dot_small(Graw_.p, G2_.p, k_ * k_, small_.p + 1, st_);
    } else if (mc_ > 0) {
        // this rank's columns of A against the whole W and its columns of H
        if (storage_ == Storage::F64) launch_error<double>(A64_.p, Wt_.p, H_.p + (size_t)k_ * c0_, k_, n_, mc_, red_part_.p, small_.p, st_);
        else if (tc_error) {
            const int kp = error_tc_kp(k_);
            ew_hi_.ensure((size_t)n_ * kp); ew_lo_.ensure((size_t)n_ * kp); rsw_.ensure(4);
            eh_hi_.ensure((size_t)mc_ * kp); eh_lo_.ensure((size_t)mc_ * kp); rsh_.ensure(4);
            unsigned long long* mxb = reinterpret_cast<unsigned long long*>(small_.p + 10);
            launch_split_rows(Wt_.p, k_, n_, ew_hi_.p, ew_lo_.p, rsw_.p, mxb, st_);
            launch_split_rows(H_.p + (size_t)k_ * c0_, k_, mc_, eh_hi_.p, eh_lo_.p, rsh_.p, mxb + 1, st_);
            launch_error_tc(A32_.p, n_, mc_, k_, ew_hi_.p, ew_lo_.p, rsw_.p, eh_hi_.p, eh_lo_.p, rsh_.p, red_part_.p, small_.p, st_);
        } else launch_error<float>(A32_.p, Wt_.p, H_.p + (size_t)k_ * c0_, k_, n_, mc_, red_part_.p, small_.p, st_);
    }
    if (comm_ && !identity) comm_->allreduce_sum_f64(small_.p, 2, st_);
    launch_factor_stats(Wt_.p, k_, n_, red_part_.p, small_.p + 2, st_);
    launch_factor_stats(H_.p, k_, m_, red_part_.p, small_.p + 5, st_);
    timer.end(st_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(host_small_.p, small_.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    sync();
    d2h_bytes += 8 * sizeof(double);
    if (identity) {
        out->sum_sq = sum_sq_a_ - 2.0 * host_small_.p[0] + host_small_.p[1];
        out->sum_kl = std::nan("");
    } else {
        out->sum_sq = host_small_.p[0];
        out->sum_kl = host_small_.p[1] - (tc_error ? kl_const_sum_ : 0.0);
    }
    last_identity_ = identity;
    for (int i = 0; i < 3; i++) { out->w_stats[i] = host_small_.p[2 + i]; out->h_stats[i] = host_small_.p[5 + i]; }
}

uint64_t Engine::take_sweeps()
{
    DeviceGuard g(device_);
    unsigned long long* hp = reinterpret_cast<unsigned long long*>(host_small_.p + 12);
    if (comm_) comm_->allreduce_sum_u64(sweeps_.p, 1, st_);
    NNLM_CUDA_CHECK(cudaMemcpyAsync(hp, sweeps_.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st_));
    NNLM_CUDA_CHECK(cudaMemsetAsync(sweeps_.p, 0, sizeof(unsigned long long), st_));
    sync();
    d2h_bytes += sizeof(unsigned long long);
    return *hp;
}

}  // namespace nnlm
