// solve_kl.cu — K7/K8: the per-column solvers for the KL-divergence loss.
//   method 3: sequential quadratic-approximation coordinate descent, reference src/base_algorithms.cpp:71-116
//   method 4: Lee-Seung multiplicative rule applied coordinate after coordinate, src/base_algorithms.cpp:119-151
// and their NA variants (src/update_with_missing.cpp:119-131: the reference gathers Wt.cols(non_missing),
// A.elem(j*n + non_missing) and sum(Wt.cols(non_missing), 1); here the same sums simply skip non-finite entries).
//
// These updates do not reduce to a dense contraction: after every coordinate c the len-vector wh = Wt' h is rank-1
// updated and the ratio A/(wh+eps) is re-formed (base_algorithms.cpp:141-143), so the work per column is
// k x (one division + two FMAs) x len, strictly sequential in c. One CTA owns one column; thread t owns the entries
// i = t, t+NT, ... of the column (A_j and wh are touched with unit stride across the CTA). The fixed factor is read
// through its row-major copy Yr[c][i] (i contiguous) so that the pass for coordinate c is one coalesced stream; it is
// shared by all columns and stays L2-resident (k*len*8 bytes). wh lives in registers when len <= NT*WHR, otherwise
// in a per-CTA global scratch row. The rank-1 update of coordinate c is fused into the pass of coordinate c+1.
// Sums over i are reduced in a fixed order (thread-strided partials -> warp butterfly -> 8 warp totals in order),
// so results are reproducible run to run; they differ from the reference's BLAS/Armadillo order by rounding only.
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int WHR = 8;          // register-resident wh entries per thread (len <= 2048)

struct Red2 { double a, b; };

// fixed-order CTA reduction of two doubles; result broadcast to every thread. `buf` is 2*NW doubles of shared memory.
__device__ __forceinline__ Red2 block_sum2(double a, double b, double* buf)
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();                       // previous readers of buf are done
    if (lane == 0) { buf[w] = a; buf[NW + w] = b; }
    __syncthreads();
    Red2 r{0.0, 0.0};
#pragma unroll
    for (int x = 0; x < NW; x++) { r.a += buf[x]; r.b += buf[NW + x]; }
    return r;
}

template <typename TA> __device__ __forceinline__ bool missing_val(TA v);
template <> __device__ __forceinline__ bool missing_val<double>(double v) { return is_missing(v); }
template <> __device__ __forceinline__ bool missing_val<float>(float v) { return ((__float_as_uint(v) >> 23) & 0xffu) == 0xffu; }

// METHOD 3 or 4. REGS: wh held in registers.
template <int METHOD, typename TA, bool REGS>
__global__ void __launch_bounds__(NT)
k_solve_kl(double* __restrict__ X, const double* __restrict__ Yr, const TA* __restrict__ A, const double* __restrict__ sumY,
           const uint8_t* __restrict__ mask, int k, int64_t len, int64_t ncol, double b0, double b1, double b2,
           unsigned max_iter, double rel_tol, int with_missing, double* __restrict__ wh_scratch,
           unsigned long long* __restrict__ sweeps)
{
    extern __shared__ double sm[];
    double* hs = sm;                 // [k] current column of X
    double* sw = sm + k;             // [k] sum of Y over the non-missing entries of this column
    double* red = sm + 2 * k;        // [2*NW]
    __shared__ int s_allmasked;

    double whr[REGS ? WHR : 1];
    double* whg = REGS ? nullptr : wh_scratch + (int64_t)blockIdx.x * len;

    for (int64_t col = blockIdx.x; col < ncol; col += gridDim.x) {
        const TA* Aj = A + len * col;
        const uint8_t* mcol = mask ? mask + (int64_t)k * col : nullptr;
        __syncthreads();
        if (threadIdx.x == 0) s_allmasked = 1;
        __syncthreads();
        for (int c = threadIdx.x; c < k; c += NT) {
            hs[c] = X[c + (int64_t)k * col];
            sw[c] = sumY[c];
            if (!(mcol && mcol[c])) s_allmasked = 0;       // benign race: all writers store 0
        }
        __syncthreads();
        if (mcol && s_allmasked) continue;                  // src/update_with_missing.cpp:33-34,77-78

        double sumH = 0.0;
        for (int c = 0; c < k; c++) sumH += hs[c];          // sum(Hj), same order on every thread

        // visit every owned entry i with a reference to its wh slot (registers or the scratch row)
        auto for_owned = [&](auto&& f) {
            if (REGS) {
#pragma unroll
                for (int r = 0; r < WHR; r++) {
                    const int64_t i = threadIdx.x + (int64_t)r * NT;
                    if (i < len) f(i, whr[r]);
                }
            } else {
                for (int64_t i = threadIdx.x; i < len; i += NT) f(i, whg[i]);
            }
        };

        // wh = Yr' h  (base_algorithms.cpp:82,133)
        for_owned([&](int64_t i, double& w) {
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < k; c++) s = fma(Yr[(int64_t)c * len + i], hs[c], s);
            w = s;
        });

        unsigned t = 0;
        bool cont = true;
        int pend_c = -1;            // coordinate whose rank-1 update of wh is still pending
        double pend_d = 0.0;
        for (; t < max_iter && cont; t++) {
            bool flag = false;
            for (int c = 0; c < k; c++) {
                if (mcol && mcol[c]) continue;
                const double* yc = Yr + (int64_t)c * len;
                const double* yp = Yr + (int64_t)(pend_c < 0 ? 0 : pend_c) * len;
                double pa = 0.0, pb = 0.0, ps = 0.0;
                const bool need_sw = with_missing && t == 0;
                const int pc = pend_c;
                auto entry = [&](double w, double ypv, TA av, double y, double& wout) {
                    if (pc >= 0) w = fma(pend_d, ypv, w);
                    wout = w;
                    if (with_missing && missing_val<TA>(av)) return;
                    const double a = static_cast<double>(av);
                    if (METHOD == 3) {
                        const double mu = y / (w + TINY_NUM);               // :97
                        pa = fma(a, mu * mu, pa);                             // dot(Aj, square(mu))
                        pb = fma(a, mu, pb);
                    } else {
                        pa = fma(y, a / (w + TINY_NUM), pa);                 // :141
                    }
                    if (need_sw) ps += y;
                };
                if (REGS) {
#pragma unroll
                    for (int r = 0; r < WHR; r++) {
                        const int64_t i = threadIdx.x + (int64_t)r * NT;
                        if (i < len) entry(whr[r], pc >= 0 ? yp[i] : 0.0, Aj[i], yc[i], whr[r]);
                    }
                } else {
                    // four entries per thread and trip, every load issued before the first use: the rolled loop paid the
                    // L2 latency of its four loads once per entry (memory-latency-bound at 8 warps per scheduler)
                    constexpr int UNR = 4;
                    for (int64_t i0 = threadIdx.x; i0 < len; i0 += (int64_t)NT * UNR) {
                        double w[UNR], ypv[UNR], y[UNR];
                        TA av[UNR];
#pragma unroll
                        for (int u = 0; u < UNR; u++) {
                            const int64_t i = i0 + (int64_t)u * NT;
                            const bool in = i < len;
                            w[u] = in ? whg[i] : 0.0;
                            ypv[u] = (in && pc >= 0) ? yp[i] : 0.0;
                            av[u] = in ? Aj[i] : TA(0);
                            y[u] = in ? yc[i] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < UNR; u++) {
                            const int64_t i = i0 + (int64_t)u * NT;
                            if (i < len) {
                                double wout;
                                entry(w[u], ypv[u], av[u], y[u], wout);
                                if (pc >= 0) whg[i] = wout;
                            }
                        }
                    }
                }
                pend_c = -1;
                if (need_sw) {
                    const Red2 r2 = block_sum2(ps, 0.0, red);
                    if (threadIdx.x == 0) sw[c] = r2.a;
                    __syncthreads();
                }
                const Red2 rs = block_sum2(pa, pb, red);
                const double hc = hs[c];
                if (METHOD == 3) {
                    double a2 = rs.a, b = rs.b - sw[c];
                    a2 += b0;                                                 // :100 (before a*h, as in the code)
                    b += a2 * hc - b2 - b1 * (sumH - hc);
                    double cand = b / (a2 + TINY_NUM);
                    if (cand < 0) cand = 0;
                    if (cand != hc) {
                        pend_c = c; pend_d = cand - hc;
                        const double e = 2 * fabs(hc - cand) / (cand + hc + TINY_NUM);
                        flag = flag || (e > rel_tol);
                        sumH += cand - hc;
                        __syncthreads();
                        if (threadIdx.x == 0) hs[c] = cand;
                        __syncthreads();
                    }
                } else {
                    const double ratio = rs.a / (sw[c] + b0 * hc + b1 * (sumH - hc) + b2);
                    const double step = (ratio - 1) * hc;
                    pend_c = c; pend_d = step;
                    sumH += step;
                    __syncthreads();
                    if (threadIdx.x == 0) hs[c] = hc * ratio;
                    __syncthreads();
                    const double e = 2 * fabs(ratio - 1) / (ratio + 1);
                    flag = flag || (e > rel_tol);
                }
            }
            cont = flag || (0.0 > rel_tol);
        }
        // (a pending rank-1 update only concerns wh, which is discarded here)
        __syncthreads();
        for (int c = threadIdx.x; c < k; c += NT) X[c + (int64_t)k * col] = hs[c];
        if (threadIdx.x == 0 && t) atomicAdd(sweeps, (unsigned long long)t);
    }
}

template <int METHOD, typename TA>
void launch_m(double* X, const double* Yr, const TA* A, const double* sumY, const uint8_t* mask, int k, int64_t len,
              int64_t ncol, const double* pen, unsigned max_iter, double rel_tol, int with_missing, double* wh_scratch,
              int grid, unsigned long long* sweeps, cudaStream_t st)
{
    const size_t smem = sizeof(double) * (2 * (size_t)k + 2 * NW);
    if (len <= (int64_t)NT * WHR)
        k_solve_kl<METHOD, TA, true><<<grid, NT, smem, st>>>(X, Yr, A, sumY, mask, k, len, ncol, pen[0], pen[1], pen[2],
                                                             max_iter, rel_tol, with_missing, nullptr, sweeps);
    else
        k_solve_kl<METHOD, TA, false><<<grid, NT, smem, st>>>(X, Yr, A, sumY, mask, k, len, ncol, pen[0], pen[1], pen[2],
                                                              max_iter, rel_tol, with_missing, wh_scratch, sweeps);
    NNLM_LAUNCHED();
}

}  // namespace

int solve_kl_grid(int64_t ncol)
{
    return (int)std::max<int64_t>(1, std::min<int64_t>(ncol, 148 * 8));
}

size_t solve_kl_scratch_doubles(int64_t len, int64_t ncol)
{
    if (len <= (int64_t)NT * WHR) return 0;
    return (size_t)solve_kl_grid(ncol) * (size_t)len;
}

template <typename TA>
void launch_solve_kl(int method, double* X, const double* Yr, const TA* A, const double* sumY, const uint8_t* mask, int k,
                     int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol, int with_missing,
                     double* wh_scratch, unsigned long long* sweeps, cudaStream_t st)
{
    NNLM_REQUIRE(method == 3 || method == 4, "solve_kl handles methods 3 and 4");
    NNLM_REQUIRE(k >= 1 && k <= 2048, "rank k must be in [1, 2048] for the KL solvers");
    if (ncol <= 0) return;
    const int grid = solve_kl_grid(ncol);
    if (method == 3) launch_m<3, TA>(X, Yr, A, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, with_missing, wh_scratch, grid, sweeps, st);
    else             launch_m<4, TA>(X, Yr, A, sumY, mask, k, len, ncol, pen, max_iter, rel_tol, with_missing, wh_scratch, grid, sweeps, st);
}
template void launch_solve_kl<double>(int, double*, const double*, const double*, const double*, const uint8_t*, int, int64_t,
                                      int64_t, const double*, unsigned, double, int, double*, unsigned long long*, cudaStream_t);
template void launch_solve_kl<float>(int, double*, const double*, const float*, const double*, const uint8_t*, int, int64_t,
                                     int64_t, const double*, unsigned, double, int, double*, unsigned long long*, cudaStream_t);

}  // namespace nnlm
