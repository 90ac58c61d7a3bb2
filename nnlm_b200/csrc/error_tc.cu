// error_tc.cu — a8 on the tensor cores: the loss bookkeeping of the outer loop (src/nnmf.cpp:121-141, 164-177) for the
// fast-precision path. The reference materialises Ahat = W.t()*H and makes two more passes over it (one with a log) at
// every `trace` iteration — every SECOND iteration with R's default trace = 100 / inner.max.iter = 2 (R/nnmf.R:138). The
// fp64 kernel of error_eval.cu never materialises Ahat but spends 2 n m k fp64 FMAs + n m fp64 logs (4.8 ms at config 2,
// more than two whole ANLS iterations). Here a 128 x 128 tile of Ahat is ONE tcgen05 contraction over the rank
// (K = 64 or 128) in TMEM and the epilogue folds it against the matching tile of A straight out of TMEM:
//   * operands: W and H as fp16 hi/lo planes [row][K] (K-major: column i of the k x n factor is already contiguous), each
//     factor scaled by one power of two, Ahat = (hi.hi + 2^-11 (hi.lo + lo.hi)) / (s_W s_H) — the scheme of cross_tc.cu;
//   * square loss: r = a - Ahat in fp32 (Ahat carries ~2^-22 relative error, random in sign: 1e-6 of a residual of 0.1);
//   * KL: the reference adds a constant term and a variable term that cancel to ~1e-5 of their size. Written per entry,
//       [(a+e) log(a+e) - a] + [-(a+e) log(Ahat+e) + Ahat] = (a+e) * g(x),  x = (Ahat - a) / (a+e),  g(x) = x - log1p(x) >= 0,
//     a sum of non-negative terms with no cancellation, which fp32 evaluates to ~1e-7 relative (series for |x| < 1/8);
//     the kernel returns sum (a+e) g(x); the engine subtracts the constant term so callers keep adding it like the reference;
//   * per-thread partial sums in fp32 over 32 entries, then fp64; CTA partials reduced in a fixed order.
// HBM-bound on the one pass over the fp32 copy of A (2 GB at config 2). Missing entries (non-finite a) are skipped as
// in src/nnmf.cpp:124-125. Warp roles as cross_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue.
#include <algorithm>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace nnlm {

using namespace tc;

namespace {

constexpr int TM = 128;            // rows of W (entries i) per tile
constexpr int TN = 128;            // rows of H (columns j) per tile
constexpr int TK = 64;             // rank elements per k-block (128 bytes of fp16)
constexpr int E_EPI_WARPS = 16;         // four per TMEM lane quarter, 32 columns of the tile each
constexpr int E_THREADS = 32 * (2 + E_EPI_WARPS);
constexpr int E_STAGES = 3;
constexpr int PLANE_BYTES = TM * TK * 2;                 // 16 KB (TM == TN)
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;             // W hi, W lo, H hi, H lo

struct ErrParams {
    int64_t n, m;              // rows of A, columns of A held by this rank
    int64_t tiles_i, tiles_j;
    int kblocks;               // ceil(k / 64)
    const float* A;            // n x m column-major, non-finite = missing
    const float* rsw;          // [1]  1 / s_W (one power-of-two scale per factor: the fp16 halves are floating-point, so
    const float* rsh;          // [1]  1 / s_H  every entry keeps its 22 bits down to 2^-27 of the largest one)
    double* part;              // [gridDim.x][2]
    float* out;                // MODE 1: the product tile goes here, out[i + n * j] (same layout as A)
};

// g(x) = x - log1p(x) for |x| < 1/8 as x^2 (1/2 - x/3 + x^2/4 - ...): nine terms leave < 2e-9 relative
__device__ __forceinline__ float g_series(float x)
{
    float p = 1.0f / 10.0f;
    p = fmaf(-p, x, 1.0f / 9.0f);
    p = fmaf(-p, x, 1.0f / 8.0f);
    p = fmaf(-p, x, 1.0f / 7.0f);
    p = fmaf(-p, x, 1.0f / 6.0f);
    p = fmaf(-p, x, 1.0f / 5.0f);
    p = fmaf(-p, x, 1.0f / 4.0f);
    p = fmaf(-p, x, 1.0f / 3.0f);
    p = fmaf(-p, x, 1.0f / 2.0f);
    return x * x * p;
}

// (a+e) g(x) for any x > -1. Ahat ~ 0 against a > 0 (x -> -1, beyond fp32's resolution of 1 + x) is written as
// (a+e) log((a+e) / (Ahat+e)) + Ahat - a, which has no cancellation there.
__device__ __noinline__ float kl_term_general(float a, float ah)
{
    const float ae = a + 1e-16f;
    const float x = (ah - a) / ae;
    if (x < -0.9999f) return fmaf(ae, logf(ae) - logf(ah + 1e-16f), ah - a);
    return ae * (x - log1pf(x));
}

// MODE 0: the two losses. MODE 1: no loss, the tile of Ahat itself is written out in fp32 (the KL solvers start every column
// from wh = W h: formed here at tensor-core speed instead of k passes over the factor rows per column group, solve_kl_fast.cuh).
template <int MODE>
__global__ void __launch_bounds__(E_THREADS, 1)
k_error_tc(const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
           const __grid_constant__ CUtensorMap mapH_hi, const __grid_constant__ CUtensorMap mapH_lo, const ErrParams p)
{
    constexpr uint32_t TMEM_COLS = 512;                  // 2 buffers x (d0 | d1) x 128 columns
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);   // f16 x f16 -> f32, K-major

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + E_STAGES * STAGE_BYTES);
    uint64_t* empty = full + E_STAGES;
    uint64_t* tfull = empty + E_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    double* red = reinterpret_cast<double*>(tmem_slot + 2);          // [E_EPI_WARPS][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = p.tiles_i * p.tiles_j;

    if (threadIdx.x == 0) {
        for (int s = 0; s < E_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], E_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer: one stage per (tile, k-block) =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int r0 = (int)((t % p.tiles_i) * TM), c0 = (int)((t / p.tiles_i) * TN);
                for (int kb = 0; kb < p.kblocks; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(sa, &mapW_hi, &full[stage], kb * TK, r0);
                    tma_load_2d(sa + PLANE_BYTES, &mapW_lo, &full[stage], kb * TK, r0);
                    tma_load_2d(sa + 2 * PLANE_BYTES, &mapH_hi, &full[stage], kb * TK, c0);
                    tma_load_2d(sa + 3 * PLANE_BYTES, &mapH_lo, &full[stage], kb * TK, c0);
                    if (++stage == E_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
                const uint32_t buf = it & 1, tph = (it >> 1) & 1;
                mbar_wait(&tempty[buf], tph ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + buf * (2 * TN), d1 = d0 + TN;
                for (int kb = 0; kb < p.kblocks; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
                    const uint64_t w_hi = make_desc(sa), w_lo = make_desc(sa + PLANE_BYTES);
                    const uint64_t h_hi = make_desc(sa + 2 * PLANE_BYTES), h_lo = make_desc(sa + 3 * PLANE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < TK / 16; ks++) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        const uint32_t acc = (kb == 0 && ks == 0) ? 0u : 1u;
                        umma_f16(d0, w_hi + adv, h_hi + adv, IDESC, acc);
                        umma_f16(d1, w_hi + adv, h_lo + adv, IDESC, acc);
                        umma_f16(d1, w_lo + adv, h_hi + adv, IDESC, 1u);
                    }
                    tc_commit(&empty[stage]);
                    if (++stage == E_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[buf]);
            }
        }
    } else {
        // ===================================== epilogue: TMEM tile of Ahat against the tile of A =====================================
        const int ew = warp - 2;
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int part = ew >> 2;                      // which 32 columns of the tile
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        double acc_sq = 0.0, acc_kl = 0.0;
        const float unscale = p.rsw[0] * p.rsh[0];
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
            const int64_t i = (t % p.tiles_i) * TM + quarter * 32 + lane;
            const int64_t j0 = (t / p.tiles_i) * TN + part * 32;
            const uint32_t buf = it & 1, tph = (it >> 1) & 1;
            const bool row_ok = i < p.n;
            if (MODE == 1) {
                mbar_wait(&tfull[buf], tph);
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    uint32_t r0[16], r1[16];
                    const uint32_t t0 = tmem_base + lane_addr + buf * (2 * TN) + part * 32 + ch * 16;
                    TmemLd<16>::ld(t0, r0);
                    TmemLd<16>::ld(t0 + TN, r1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (ch == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[buf]);
                    }
#pragma unroll
                    for (int c = 0; c < 16; c++) {
                        const int64_t j = j0 + ch * 16 + c;
                        if (row_ok && j < p.m)
                            p.out[i + p.n * j] = fmaf(__uint_as_float(r1[c]), (float)LO_UNSCALE, __uint_as_float(r0[c])) * unscale;
                    }
                }
                continue;
            }
            // this thread's 32 entries of A: every load is issued before the tile is waited for, so
            // their latency hides behind the MMA (a first version loaded inside the compute loop, behind its branches: 64
            // serialised L2 round trips per tile, 8 ms per evaluation instead of 0.5)
            float av[32];
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const int64_t j = j0 + c;
                const bool ok = row_ok && j < p.m;
                av[c] = ok ? p.A[i + p.n * j] : __int_as_float(0x7fc00000);        // NaN = skipped like a missing entry
            }
            mbar_wait(&tfull[buf], tph);
            tc_fence_after();
            float sq = 0.0f, kl = 0.0f;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {           // two sub-chunks of 16 columns keep the TMEM read-out at 32 registers
                uint32_t r0[16], r1[16];
                const uint32_t t0 = tmem_base + lane_addr + buf * (2 * TN) + part * 32 + ch * 16;
                TmemLd<16>::ld(t0, r0);
                TmemLd<16>::ld(t0 + TN, r1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf]);          // the MMA of the tile after next may overwrite this buffer
                }
                // branch-free main path (the series form of g; NaN from a missing entry flows through and is dropped by the
                // final select), four independent accumulator pairs; entries outside the series' range are rare once the fit
                // has started and are redone per thread with the general formula
                float s4[4] = {0.f, 0.f, 0.f, 0.f}, k4[4] = {0.f, 0.f, 0.f, 0.f};
                unsigned redo = 0;
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const float a = av[ch * 16 + c];
                    const bool fin = ((__float_as_uint(a) >> 23) & 0xffu) != 0xffu;            // finite: src/nnmf.cpp:124-125
                    const float ah = fmaf(__uint_as_float(r1[c]), (float)LO_UNSCALE, __uint_as_float(r0[c])) * unscale;
                    const float r = a - ah;
                    const float ae = a + 1e-16f;
                    const float x = __fdividef(ah - a, ae);
                    const bool in_range = fabsf(x) < 0.125f;
                    redo |= (fin && !in_range) ? (1u << c) : 0u;
                    s4[c & 3] += fin ? r * r : 0.0f;
                    k4[c & 3] += (fin && in_range) ? ae * g_series(x) : 0.0f;
                }
                if (redo) {
#pragma unroll
                    for (int c = 0; c < 16; c++)
                        if ((redo >> c) & 1u)
                            k4[0] += kl_term_general(av[ch * 16 + c],
                                                     fmaf(__uint_as_float(r1[c]), (float)LO_UNSCALE, __uint_as_float(r0[c])) * unscale);
                }
                sq += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                kl += (k4[0] + k4[1]) + (k4[2] + k4[3]);
            }
            acc_sq += (double)sq;
            acc_kl += (double)kl;
        }
        acc_sq = warp_sum(acc_sq);
        acc_kl = warp_sum(acc_kl);
        if (lane == 0) { red[2 * ew] = acc_sq; red[2 * ew + 1] = acc_kl; }
    }
    tc_fence_before();
    __syncthreads();
    if (MODE == 0 && threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int w = 0; w < E_EPI_WARPS; w++) { a0 += red[2 * w]; a1 += red[2 * w + 1]; }
        p.part[2 * blockIdx.x] = a0;
        p.part[2 * blockIdx.x + 1] = a1;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// max |X| as a bit pattern (one atomicMax per block)
__global__ void __launch_bounds__(256)
k_absmax_bits(const double* __restrict__ X, int64_t total, unsigned long long* __restrict__ out)
{
    unsigned long long m = 0;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const double v = fabs(X[e]);
        if (!is_missing(v)) m = max(m, (unsigned long long)__double_as_longlong(v));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// X (k x cols, column-major fp64) -> planes [cols][KP] (row i = column i of X, zero beyond k), scaled by the power of two that
// brings max |X| into [2^13, 2^14); rs[0] = 1 / scale
__global__ void __launch_bounds__(256)
k_split_rows(const double* __restrict__ X, int k, int64_t cols, int kp, const unsigned long long* __restrict__ maxbits,
             __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ rs)
{
    const double sc = pow2_scale(__longlong_as_double((long long)maxbits[0]));
    if (blockIdx.x == 0 && threadIdx.x == 0) rs[0] = (float)(1.0 / sc);
    const int64_t total = cols * kp;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t row = e / kp;
        const int c = (int)(e % kp);
        __half h = __float2half_rn(0.f), l = h;
        if (c < k) split2(X[c + (int64_t)k * row] * sc, h, l);
        hi[e] = h;
        lo[e] = l;
    }
}

}  // namespace

bool error_tc_supported(int k) { return k >= 1 && k <= 128; }
int error_tc_kp(int k) { return k <= 64 ? 64 : 128; }
int error_tc_grid(int64_t n, int64_t m) { return (int)std::min<int64_t>(sm_count(), ceil_div(n, TM) * ceil_div(m, TN)); }

void launch_split_rows(const double* X, int k, int64_t cols, __half* hi, __half* lo, float* rs, unsigned long long* maxbits,
                       cudaStream_t st)
{
    if (cols <= 0) return;
    const int64_t total = (int64_t)k * cols;
    NNLM_CUDA_CHECK(cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long), st));
    k_absmax_bits<<<(int)std::min<int64_t>(ceil_div(total, 2048), 148 * 4), 256, 0, st>>>(X, total, maxbits);
    NNLM_LAUNCHED();
    const int kp = error_tc_kp(k);
    k_split_rows<<<(int)std::min<int64_t>(ceil_div(cols * kp, 1024), 148 * 8), 256, 0, st>>>(X, k, cols, kp, maxbits, hi, lo, rs);
    NNLM_LAUNCHED();
}

// out[0] = sum over finite a of (a - Ahat)^2; out[1] = sum over finite a of (a+e) g(x) = the reference's constant KL term
// sum[(a+e) log(a+e) - a] (src/nnmf.cpp:70-73) PLUS its variable term sum[-(a+e) log(Ahat+e) + Ahat] (:125,139): the caller
// subtracts the constant it already holds. A: n x m fp32 column-major; W planes [n][kp], H planes [m][kp] (launch_split_rows);
// part: 2 * error_tc_grid doubles.
void launch_error_tc(const float* A, int64_t n, int64_t m, int k, const __half* w_hi, const __half* w_lo, const float* rsw,
                     const __half* h_hi, const __half* h_lo, const float* rsh, double* part, double* out, cudaStream_t st)
{
    NNLM_REQUIRE(error_tc_supported(k), "tensor-core error evaluation supports rank k <= 128");
    const int kp = error_tc_kp(k);
    constexpr size_t smem = (size_t)E_STAGES * STAGE_BYTES + 1024 + 512;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(k_error_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ErrParams p;
    p.n = n; p.m = m; p.tiles_i = ceil_div(n, TM); p.tiles_j = ceil_div(m, TN); p.kblocks = kp / TK;
    p.A = A; p.rsw = rsw; p.rsh = rsh; p.part = part; p.out = nullptr;
    const CUtensorMap mW_hi = make_map(w_hi, kp, n, kp, TM), mW_lo = make_map(w_lo, kp, n, kp, TM);
    const CUtensorMap mH_hi = make_map(h_hi, kp, m, kp, TN), mH_lo = make_map(h_lo, kp, m, kp, TN);
    const int grid = error_tc_grid(n, m);
    k_error_tc<0><<<grid, E_THREADS, smem, st>>>(mW_hi, mW_lo, mH_hi, mH_lo, p);
    NNLM_LAUNCHED();
    launch_reduce_partials(part, grid, 2, out, st);
}

// out[i + n * j] = sum_c W[c, i] H[c, j] in fp32 (~2^-22 relative): W planes [n][kp], H planes [m][kp] from launch_split_rows
void launch_product_tc(int64_t n, int64_t m, int k, const __half* w_hi, const __half* w_lo, const float* rsw, const __half* h_hi,
                       const __half* h_lo, const float* rsh, float* out, cudaStream_t st)
{
    NNLM_REQUIRE(error_tc_supported(k), "the tensor-core factor product supports rank k <= 128");
    if (n <= 0 || m <= 0) return;
    const int kp = error_tc_kp(k);
    constexpr size_t smem = (size_t)E_STAGES * STAGE_BYTES + 1024 + 512;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(k_error_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ErrParams p;
    p.n = n; p.m = m; p.tiles_i = ceil_div(n, TM); p.tiles_j = ceil_div(m, TN); p.kblocks = kp / TK;
    p.A = nullptr; p.rsw = rsw; p.rsh = rsh; p.part = nullptr; p.out = out;
    const CUtensorMap mW_hi = make_map(w_hi, kp, n, kp, TM), mW_lo = make_map(w_lo, kp, n, kp, TM);
    const CUtensorMap mH_hi = make_map(h_hi, kp, m, kp, TN), mH_lo = make_map(h_lo, kp, m, kp, TN);
    k_error_tc<1><<<error_tc_grid(n, m), E_THREADS, smem, st>>>(mW_hi, mW_lo, mH_hi, mH_lo, p);
    NNLM_LAUNCHED();
}

}  // namespace nnlm
