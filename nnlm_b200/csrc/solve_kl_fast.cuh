// solve_kl_fast.cuh — K7/K8 restructured for the fast-precision path (dense A stored in fp32): the per-column KL solvers of
// src/base_algorithms.cpp:71-116 (method 3) and :119-151 (method 4) with the column state kept ON CHIP.
//
// Why. After every coordinate c the len-vector wh = Wt' h takes a rank-1 update and the ratios A/(wh+eps) are re-formed
// (base_algorithms.cpp:141-143): k sequential passes over len entries per column and sweep. The first kernel
// (solve_kl.cu) kept wh in a global scratch row and walked it, A_j and two factor rows out of L2 on every pass: 36 bytes
// of L2 traffic per entry and coordinate, 900 GB per half-iteration at config 3 — L2-bound at ~100 ms. Here
//   * a thread-block CLUSTER of S CTAs owns KLF_J columns at a time; the len entries are dealt over the S x 512 threads
//     (entry i = e*S*512 + rank*512 + tid, so every factor-row load is one coalesced stream), each thread holding its
//     E entries of wh for all KLF_J columns in REGISTERS (fp32) and the matching entries of A in SHARED memory (fp32);
//   * the only traffic per coordinate is the factor row itself (fp32 copy, len*4 bytes per cluster, shared by the KLF_J
//     columns): 25 GB per half at config 3 instead of 900. Each thread stages its entries of the NEXT row into shared memory
//     with cp.async right after a pass, so the L2 latency hides behind the reduction of the current coordinate; the row
//     stays there one more step for the pending rank-1 update (two thread-private buffers, no CTA barrier needed);
//   * ratios are formed in fp32 (MUFU reciprocal + one Newton step), per-thread partial sums in fp32 over <= 16 terms,
//     everything across threads, CTAs and coordinates (sums, h, sum(h), the update formulas) in fp64 and in a fixed order:
//     warp butterfly -> warp totals in order -> CTA records exchanged through distributed shared memory -> summed in rank
//     order by every CTA, so all CTAs of a cluster take bit-identical decisions and results are reproducible run to run.
// No cluster barrier per coordinate: the records travel as st.async stores that signal the receiver's mbarrier (exchange
// buffers and barriers alternate, so a fast CTA never overwrites a record a slow one still reads). Columns of a group advance in lock step; a column that has converged (or is fully masked) is frozen.
// Measured error against the fp64 oracle: tests/test_gpu_scale_parity.py (<= 1e-5 bar, ~1e-7 observed).
// The exact path (fp64 A), missing values and len > 65536 stay on solve_kl.cu.
#include <cooperative_groups.h>

#include <algorithm>

#pragma once
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace nnlm {

namespace klf {

constexpr int KLF_NT = 512;          // threads per CTA
constexpr int KLF_NW = KLF_NT / 32;
constexpr int KLF_J = 4;             // columns a cluster advances together
constexpr int KLF_MAXS = 8;          // portable cluster size
constexpr int KLF_NV = 2 * KLF_J;    // reduced values per coordinate: a per column, then b per column (method 3)
static_assert(KLF_J == 4, "the A tile is staged as one float4 per entry");

// a / w through the MUFU reciprocal (relative error ~2^-23, the size of one fp32 rounding: no Newton step — the pass is
// bound by issue slots, and parity holds, tests/test_gpu_scale_parity.py)
__device__ __forceinline__ float fast_div(float a, float w)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
    return a * r;
}

__device__ __forceinline__ float fast_div_newton(float a, float w)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
    r = fmaf(fmaf(-w, r, 1.0f), r, r);
    const float q = a * r;
    return fmaf(fmaf(-w, q, a), r, q);           // correctly rounded quotient up to one ulp
}

// sums of NV per-lane values over the warp with NV - 1 + log2(32 / NV) shuffles instead of 5 NV: lanes trade halves of their
// value vector on the way down (fp32; the totals cross warps, CTAs and coordinates in fp64). Result for value x ends in lane
// (x * 32 / NV) ... returned as: every lane holds the total of value (lane / (32 / NV)).
template <int NV>
__device__ __forceinline__ float warp_multi_sum(float (&v)[NV], int lane)
{
    static_assert(NV == 4 || NV == 8, "4 or 8 values");
    if (NV == 8) {
        const bool up = lane & 16;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const float send = up ? v[x] : v[x + 4];
            const float keep = up ? v[x + 4] : v[x];
            v[x] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    // 4 values left in v[0..3]
    {
        const bool up = lane & 8;
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const float send = up ? v[x] : v[x + 2];
            const float keep = up ? v[x + 2] : v[x];
            v[x] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    if (NV == 4) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
    return v[0];
}

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// ---- one-sided exchange over distributed shared memory: a store that signals the destination CTA's mbarrier ----
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// number of doubles at the head of the dynamic shared memory (hs, sw, red, xch, 2 mbarriers), rounded up to an even count
__host__ __device__ constexpr size_t klf_head_doubles(int k)
{
    return (((size_t)KLF_J * k + k + KLF_NW * KLF_NV + 2 * KLF_MAXS * KLF_NV + 2) + 1) & ~(size_t)1;
}
__device__ __forceinline__ unsigned char* sm_doubles_end(unsigned char* base, int k) { return base + sizeof(double) * klf_head_doubles(k); }

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int METHOD, int E>
__global__ void __launch_bounds__(KLF_NT, 1)
k_solve_kl_fast(double* __restrict__ X, const float* __restrict__ Y32, const float* __restrict__ A, const float* __restrict__ WH0,
                const double* __restrict__ sumY, const uint8_t* __restrict__ mask, int k, int64_t len, int64_t ncol, double b0, double b1, double b2,
                unsigned max_iter, double rel_tol, unsigned long long* __restrict__ sweeps)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int64_t cid = blockIdx.x / S, nclusters = gridDim.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* hs = reinterpret_cast<double*>(smem_raw);            // [KLF_J][k] current columns of X
    double* sw = hs + (size_t)KLF_J * k;                         // [k] rowSums of the fixed factor (:27)
    double* red = sw + k;                                        // [KLF_NW][KLF_NV]
    double* xch = red + KLF_NW * KLF_NV;                         // [2][KLF_MAXS][KLF_NV] records of every CTA of the cluster
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * KLF_MAXS * KLF_NV);   // [2] one mbarrier per exchange buffer
    // (the double section is padded to an even count so the float4 tile of A below stays 16-byte aligned for every k)
    float* dS = reinterpret_cast<float*>(sm_doubles_end(smem_raw, k));   // [KLF_J] step d of the coordinate just solved; [KLF_J] = "another sweep" flag
    float* aS = dS + 2 * KLF_J;                                  // [E][KLF_J][KLF_NT] this thread's entries of A
    float* yS = aS + (size_t)E * KLF_J * KLF_NT;                 // [2][E][KLF_NT] this thread's entries of two factor rows

    const int64_t stride = (int64_t)S * KLF_NT;
    const int64_t i_first = (int64_t)rank * KLF_NT + tid;
    const double tiny = TINY_NUM;

    // asynchronous staging of this thread's entries of factor row c into buffer b (slots are thread-private: no CTA barrier).
    // Everything that does not depend on (c, b) is hoisted: a validity bit per entry, one 32-bit shared address, one
    // 64-bit global pointer; slots of entries beyond len are zeroed once below and never written again.
    unsigned valid = 0;
#pragma unroll
    for (int e = 0; e < E; e++) valid |= (i_first + (int64_t)e * stride < len ? 1u : 0u) << e;
    const uint32_t y_sbase = (uint32_t)__cvta_generic_to_shared(yS) + 4u * (uint32_t)tid;
    const float* const y_gbase = Y32 + i_first;
    const uint32_t stride_u = (uint32_t)stride;
    auto prefetch_row = [&](int c, int b) {
        const float* row = y_gbase + (int64_t)c * len;
        const uint32_t sb = y_sbase + (uint32_t)b * (uint32_t)(E * KLF_NT * 4);
#pragma unroll
        for (int e = 0; e < E; e++)
            if ((valid >> e) & 1u)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sb + (uint32_t)(e * KLF_NT * 4)), "l"(row + (size_t)((uint32_t)e * stride_u)) : "memory");
    };

    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bars), xch_s = (uint32_t)__cvta_generic_to_shared(xch);
    if (tid == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_s + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();         // every CTA's barriers exist before a peer signals them

    for (int c = tid; c < k; c += KLF_NT) sw[c] = sumY[c];
#pragma unroll
    for (int e = 0; e < 2 * E; e++) yS[(size_t)e * KLF_NT + tid] = 0.0f;       // a never-staged buffer must not hold NaN (k == 1)
#ifdef NNLM_KLF_PROF
    long long pk_wait = 0, pk_pass = 0, pk_red = 0, pk_send = 0, pk_xch = 0, pk_upd = 0, pk_n = 0, pk_ga = 0, pk_gi = 0, pk_g = 0;
#endif
    unsigned step = 0;      // coordinate steps since kernel start: parity of the cluster exchange buffers (never reset, so a CTA
                            // that runs ahead into the next column group cannot overwrite a record a peer still reads)

    for (int64_t grp = cid; grp * KLF_J < ncol; grp += nclusters) {
        const int64_t col0 = grp * KLF_J;
#ifdef NNLM_KLF_PROF
        long long pg0 = clock64();
#endif
        __syncthreads();
        for (int e2 = tid; e2 < KLF_J * k; e2 += KLF_NT) {
            const int j = e2 / k, c = e2 % k;
            hs[e2] = (col0 + j < ncol) ? X[c + (int64_t)k * (col0 + j)] : 0.0;
        }
        // this thread's entries of A for the group's columns
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int64_t i = i_first + (int64_t)e * stride;
            float a4[KLF_J];
#pragma unroll
            for (int j = 0; j < KLF_J; j++) a4[j] = (((valid >> e) & 1u) && col0 + j < ncol) ? A[i + len * (col0 + j)] : 0.0f;
            reinterpret_cast<float4*>(aS)[e * KLF_NT + tid] = make_float4(a4[0], a4[1], a4[2], a4[3]);     // one LDS.128 per entry later
        }
        __syncthreads();

        // per-column control state lives in lane j of warp 0 of every CTA (identical across the cluster by construction)
        bool active = false, cont = false, flag = false;
        double sumH = 0.0, rden = 0.0;
        unsigned tcount = 0;
        if (warp == 0 && lane < KLF_J) {
            const int j = lane;
            bool any_free = false;
            for (int c = 0; c < k; c++) {
                sumH += hs[j * k + c];                                                // sum(Hj), coordinate order
                any_free = any_free || !(mask && col0 + j < ncol && mask[c + (int64_t)k * (col0 + j)]);
            }
            active = (col0 + j < ncol) && any_free;                                  // src/update_with_missing.cpp:33-34
            cont = active;
        }

#ifdef NNLM_KLF_PROF
        long long pg1 = clock64(); pk_ga += pg1 - pg0;
#endif
        // wh = Yr' h (base_algorithms.cpp:82,133), fp32
        float wh[E][KLF_J];
#pragma unroll
        for (int e = 0; e < E; e++)
#pragma unroll
            for (int j = 0; j < KLF_J; j++) wh[e][j] = 0.0f;
        if (WH0 != nullptr) {
            // the product was formed on the tensor cores (error_tc.cu, launch_product_tc): one coalesced load per entry
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int64_t i = i_first + (int64_t)e * stride;
#pragma unroll
                for (int j = 0; j < KLF_J; j++)
                    wh[e][j] = (((valid >> e) & 1u) && col0 + j < ncol) ? __ldg(WH0 + i + len * (col0 + j)) : 0.0f;
            }
        } else
        // (rows straight from L2 into registers, two to four rows = up to 26 loads in flight per thread: staged one row ahead through shared
        // memory like the sweeps below, every coordinate paid the full L2 latency — this loop was 60 % of the H-half at config 3)
        {
            constexpr int UB = E <= 6 ? 4 : (E <= 9 ? 3 : 2);          // rows in flight: as many as the register budget takes
            int c = 0;
            for (; c + UB <= k; c += UB) {
                float y4[UB][E];
#pragma unroll
                for (int u = 0; u < UB; u++)
#pragma unroll
                    for (int e = 0; e < E; e++)
                        y4[u][e] = ((valid >> e) & 1u) ? __ldg(y_gbase + (int64_t)(c + u) * len + (size_t)((uint32_t)e * stride_u)) : 0.0f;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    float hc[KLF_J];
#pragma unroll
                    for (int j = 0; j < KLF_J; j++) hc[j] = (float)hs[j * k + c + u];
#pragma unroll
                    for (int e = 0; e < E; e++)
#pragma unroll
                        for (int j = 0; j < KLF_J; j++) wh[e][j] = fmaf(y4[u][e], hc[j], wh[e][j]);
                }
            }
            for (; c < k; c++) {
                float y1[E];
#pragma unroll
                for (int e = 0; e < E; e++)
                    y1[e] = ((valid >> e) & 1u) ? __ldg(y_gbase + (int64_t)c * len + (size_t)((uint32_t)e * stride_u)) : 0.0f;
                float hc[KLF_J];
#pragma unroll
                for (int j = 0; j < KLF_J; j++) hc[j] = (float)hs[j * k + c];
#pragma unroll
                for (int e = 0; e < E; e++)
#pragma unroll
                    for (int j = 0; j < KLF_J; j++) wh[e][j] = fmaf(y1[e], hc[j], wh[e][j]);
            }
        }

#ifdef NNLM_KLF_PROF
        pk_gi += clock64() - pg1; pk_g++;
#endif
        float dprev[KLF_J];                       // pending rank-1 step of the previous coordinate (applied with its row)
#pragma unroll
        for (int j = 0; j < KLF_J; j++) dprev[j] = 0.0f;
        unsigned ystep = 0;                       // coordinate steps done in this group: parity of the factor-row buffers
        prefetch_row(0, 0);                       // (buffer 0 was last read two rows ago when k >= 2; k == 1 re-reads row 0 anyway)
        bool more = true;

        for (unsigned t = 0; t < max_iter && more; t++) {
            for (int c = 0; c < k; c++) {
                const int cur = ystep & 1;
#ifdef NNLM_KLF_PROF
                long long pk0 = clock64();
#endif
                cp_async_wait_all();
#ifdef NNLM_KLF_PROF
                long long pk1 = clock64(); pk_wait += pk1 - pk0;
#endif
                // ---- the pass of coordinate c over this thread's entries ----
                float pa[KLF_J], pb[KLF_J];
#pragma unroll
                for (int j = 0; j < KLF_J; j++) { pa[j] = 0.0f; pb[j] = 0.0f; }
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const float y = yS[((size_t)cur * E + e) * KLF_NT + tid];
                    const float yprev = yS[((size_t)(cur ^ 1) * E + e) * KLF_NT + tid];      // row of the previous step (dprev = 0 at step 0)
                    const float4 a4 = reinterpret_cast<const float4*>(aS)[e * KLF_NT + tid];
                    const float av[KLF_J] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                    for (int j = 0; j < KLF_J; j++) {
                        const float w = fmaf(dprev[j], yprev, wh[e][j]);             // Ajt += d * Wt.row(previous)  (:106,143)
                        wh[e][j] = w;
                        const float a = av[j];
                        if (METHOD == 3) {
                            const float we = w + 1e-16f;
                            const float mu = fast_div_newton(y, we);                  // :97
                            pa[j] = fmaf(a * mu, mu, pa[j]);                          // dot(Aj, square(mu))
                            // the reference forms dot(Aj, mu) - sumW(k) (:99), two sums that cancel to the size of the gradient;
                            // with sumW(k) = sum_i y_i and y_i = mu_i (wh_i + eps) the difference is sum_i (A_i - wh_i - eps) mu_i,
                            // accumulated here entry by entry so that fp32 keeps its relative precision on the small result
                            pb[j] = fmaf(a - we, mu, pb[j]);
                        } else {
                            pa[j] = fmaf(y, fast_div(a, w + 1e-16f), pa[j]);          // dot(Wt.row(c), Aj / (wh + eps))  (:141)
                        }
                    }
                }
                // the row after this one goes into the buffer the pass has just finished reading; it lands during the reduction
                {
                    const int cn = (c + 1 < k) ? c + 1 : 0;
                    prefetch_row(cn, cur ^ 1);
                }
#ifdef NNLM_KLF_PROF
                long long pk2 = clock64(); pk_pass += pk2 - pk1;
#endif
                // ---- fixed-order reduction: warp butterfly (fp32) -> warp totals -> CTA record -> cluster records (fp64) ----
                constexpr int NVM = METHOD == 3 ? 2 * KLF_J : KLF_J;       // values reduced per coordinate: a (and b) per column
                float vals[NVM];
#pragma unroll
                for (int j = 0; j < KLF_J; j++) { vals[j] = pa[j]; if (METHOD == 3) vals[(NVM == 2 * KLF_J ? KLF_J : 0) + j] = pb[j]; }
                const float tot = warp_multi_sum<NVM>(vals, lane);
                if ((lane & 3) == 0 && lane < 4 * NVM) red[warp * KLF_NV + (lane >> 2)] = (double)tot;
                __syncthreads();
                // Each CTA pushes its record into every CTA of the cluster (itself included) with stores that complete a
                // transaction on the DESTINATION's mbarrier: no cluster-wide barrier per coordinate. Buffers and barriers
                // alternate; a peer can only send step s+2 after it has received this CTA's step s+1, which is sent after
                // this CTA has finished reading step s, so a record is never overwritten while it is still needed.
                const uint32_t buf = step & 1;
                double* mine = xch + buf * KLF_MAXS * KLF_NV;
#ifdef NNLM_KLF_PROF
                long long pk3 = clock64(); pk_red += pk3 - pk2;
                long long pk4 = pk3;
#endif
                if (warp == 0) {
                    double s = 0.0;
                    if (lane < NVM) {
#pragma unroll
                        for (int w2 = 0; w2 < KLF_NW; w2++) s += red[w2 * KLF_NV + lane];
                    }
                    if (S == 1) {                      // a single CTA: the record stays local, no distributed shared memory involved
                        if (lane < NVM) mine[lane] = s;
                        __syncwarp();
                    } else {
                        if (lane == 0) mbar_arrive_expect_tx(bar_s + 8 * buf, (uint32_t)(S * NVM * 8));
                        if (lane < NVM) {
                            const uint32_t slot = xch_s + 8u * (buf * KLF_MAXS * KLF_NV + rank * KLF_NV + lane);
                            for (int r = 0; r < S; r++) st_async_f64(map_to_rank(slot, r), s, map_to_rank(bar_s + 8 * buf, r));
                        }
#ifdef NNLM_KLF_PROF
                        pk4 = clock64(); pk_send += pk4 - pk3;
#endif
                    }
                    // while the records travel: the reciprocal of the multiplicative rule's denominator (:141-142) — it depends on
                    // h and sum(h) only, and a fp64 division is ~30 dependent instructions that used to sit behind the wait
                    if (METHOD == 4 && lane < KLF_J) {
                        const double hc = hs[lane * k + c];
                        rden = 1.0 / (sw[c] + b0 * hc + b1 * (sumH - hc) + b2);
                    }
                    if (S > 1) mbar_wait_cluster(bar_s + 8 * buf, (step >> 1) & 1);
                }
#ifdef NNLM_KLF_PROF
                long long pk5 = clock64(); pk_xch += pk5 - pk4;
#endif
                // ---- the update of coordinate c: lane j of warp 0 owns column j (every CTA computes the same numbers) ----
                if (warp == 0 && lane < KLF_J) {
                    const int j = lane;
                    double ta = 0.0, tb = 0.0;
                    for (int r = 0; r < S; r++) { ta += mine[r * KLF_NV + j]; if (METHOD == 3) tb += mine[r * KLF_NV + KLF_J + j]; }
                    const bool live = cont && !(mask && mask[c + (int64_t)k * (col0 + j)]);
                    double d = 0.0;
                    if (live) {
                        const double hc = hs[j * k + c];
                        double hn = hc;
                        if (METHOD == 3) {
                            double a2 = ta, b = tb;                                   // tb = dot(Aj, mu) - sumW(c), see the pass
                            a2 += b0;                                                 // :100 (before a*h, as in the code)
                            b += a2 * hc - b2 - b1 * (sumH - hc);
                            double cand = b / (a2 + tiny);
                            if (cand < 0) cand = 0;
                            if (cand != hc) {
                                d = cand - hc;
                                flag = flag || (2 * fabs(hc - cand) > rel_tol * (cand + hc + tiny));
                                hn = cand;
                            }
                        } else {
                            const double ratio = ta * rden;                                   // :141-142 (reciprocal formed above)
                            d = (ratio - 1) * hc;
                            hn = hc * ratio;
                            flag = flag || (2 * fabs(ratio - 1) > rel_tol * (ratio + 1));     // 2|r-1|/(r+1) > tol, r + 1 > 0
                        }
                        sumH += d;
                        hs[j * k + c] = hn;
                    }
                    dS[j] = (float)d;
                    if (c == k - 1) {                                                 // end of the sweep: who goes on
                        if (cont) tcount++;
                        cont = cont && (flag || (0.0 > rel_tol));
                        flag = false;
                        dS[KLF_J + j] = cont ? 1.0f : 0.0f;
                    }
                }
                __syncthreads();
#ifdef NNLM_KLF_PROF
                pk_upd += clock64() - pk5; pk_n++;
#endif
#pragma unroll
                for (int j = 0; j < KLF_J; j++) dprev[j] = dS[j];
                step++;
                ystep++;
            }
            more = false;
#pragma unroll
            for (int j = 0; j < KLF_J; j++) more = more || (dS[KLF_J + j] != 0.0f);
        }
        cp_async_wait_all();                      // the row prefetched after the last pass is not used
        __syncthreads();
        if (rank == 0) {
            for (int e2 = tid; e2 < KLF_J * k; e2 += KLF_NT) {
                const int j = e2 / k, c = e2 % k;
                if (col0 + j < ncol) X[c + (int64_t)k * (col0 + j)] = hs[e2];
            }
        }
        if (rank == 0 && warp == 0) {
            unsigned long long tot = active ? tcount : 0;
#pragma unroll
            for (int x = 16; x > 0; x >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, x);
            if (lane == 0 && tot) atomicAdd(sweeps, tot);
        }
    }
#ifdef NNLM_KLF_PROF
    if (tid == 0 && cid == 0 && pk_n > 0)
        printf("klf prof S=%d E=%d rank=%d steps=%lld | per step: row wait %lld pass %lld reduce %lld send %lld exchange wait %lld update %lld cycles | per group (%lld): stage A+h %lld, wh init %lld\n",
               S, E, rank, pk_n, pk_wait / pk_n, pk_pass / pk_n, pk_red / pk_n, pk_send / pk_n, pk_xch / pk_n, pk_upd / pk_n, pk_g, pk_ga / pk_g, pk_gi / pk_g);
#endif
    if (S > 1) cluster.sync();      // no CTA exits while a peer may still write into its exchange buffer
}

inline size_t klf_smem(int k, int e)
{
    return sizeof(double) * klf_head_doubles(k)
         + sizeof(float) * (2 * KLF_J + (size_t)e * KLF_J * KLF_NT + 2 * (size_t)e * KLF_NT);
}

struct KlfShape { int S, E; };
// entries per thread E = ceil(len / (S * 512)) <= 16, instantiated for every value so no pass iterates over padding.
// Smallest cluster that fits: fewer CTAs per cluster means more clusters (more column groups in flight per GPU) and a
// cheaper cluster barrier; the per-step pass time is proportional to E either way.
inline bool klf_shape(int64_t len, KlfShape* out)
{
    for (int S = 1; S <= KLF_MAXS; S *= 2) {
        const int64_t per = ceil_div(len, (int64_t)S * KLF_NT);
        if (per <= 16) { out->S = S; out->E = (int)std::max<int64_t>(per, 1); return true; }
    }
    return false;
}

template <int METHOD, int E>
void launch_e(const KlfShape& sh, double* X, const float* Y32, const float* A, const float* WH0, const double* sumY, const uint8_t* mask, int k,
              int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol, unsigned long long* sweeps,
              cudaStream_t st)
{
    auto kern = k_solve_kl_fast<METHOD, E>;
    const size_t smem = klf_smem(k, E);
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, KLF_J);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(148 / sh.S * sh.S));
    cfg.blockDim = dim3(KLF_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)sh.S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // One cluster per slot that can be RESIDENT: a cluster must sit inside one GPC, so 148 / S overestimates (18 clusters of 8
    // were launched where 15-16 fit; the rest ran as a second wave and doubled the H-half of config 3: profiles/r2_c_kl.md)
    static int resident[KLF_MAXS + 1][17] = {};
    int& res = resident[sh.S][E];
    if (res == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148 / sh.S; }
        res = n;
    }
    const int64_t clusters = std::max<int64_t>(1, std::min<int64_t>(groups, std::min<int64_t>(res, 148 / sh.S)));
    cfg.gridDim = dim3((unsigned)(clusters * sh.S));
    NNLM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, X, Y32, A, WH0, sumY, mask, k, len, ncol, pen[0], pen[1], pen[2], max_iter, rel_tol, sweeps));
    NNLM_LAUNCHED();
}

template <int METHOD, int E0>
void launch_range(const KlfShape& sh, double* X, const float* Y32, const float* A, const float* WH0, const double* sumY, const uint8_t* mask, int k,
                  int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol, unsigned long long* sweeps,
                  cudaStream_t st)
{
    // E0 .. E0 + 3
    switch (sh.E - E0) {
        case 0: launch_e<METHOD, E0>(sh, X, Y32, A, WH0, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        case 1: launch_e<METHOD, E0 + 1>(sh, X, Y32, A, WH0, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        case 2: launch_e<METHOD, E0 + 2>(sh, X, Y32, A, WH0, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        default: launch_e<METHOD, E0 + 3>(sh, X, Y32, A, WH0, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
    }
}


#define NNLM_KLF_ARGS const KlfShape& sh, double* X, const float* Y32, const float* A, const float* WH0, const double* sumY, const uint8_t* mask, int k, \
    int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st
#define NNLM_KLF_PASS sh, X, Y32, A, WH0, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st
// explicit-instantiation entry points, one translation unit each (build parallelism): method x entries-per-thread range
void launch_m3_lo(NNLM_KLF_ARGS);   // method 3, E 1..8
void launch_m3_hi(NNLM_KLF_ARGS);   // method 3, E 9..16
void launch_m4_lo(NNLM_KLF_ARGS);
void launch_m4_hi(NNLM_KLF_ARGS);

}  // namespace klf
}  // namespace nnlm
