// synth.cu — the synthetic workload of SURVEY.md §8(d) / BASELINE.md §4, generated on the device so that the benchmark
// configurations (50000 x 10000 and larger) never need a host-side generator pass:
//   u(seed, idx) = (splitmix64(seed * 0x9E3779B97F4A7C15 + idx) >> 11) * 2^-53       (counter-based, no stream state)
//   Wtrue = u(base+1) (n x k),  Htrue = u(base+2) (k x m_global),  A = Wtrue*Htrue + noise * u(base+3, i + n*j)
//   A[i,j] = NaN  iff  u(base+4, i + n*j) < na_frac
// j is the GLOBAL column index, so a rank holding columns [col0, col0+m) generates exactly its shard of the global matrix.
// tests/conftest.py::splitmix_uniform is the numpy twin of u().
#include "kernels.cuh"

namespace nnlm {

namespace {

__host__ __device__ __forceinline__ double splitmix_u(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void k_uniform(double* __restrict__ out, int64_t count, uint64_t seed, uint64_t offset, double scale)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = scale * splitmix_u(seed, offset + (uint64_t)e);
}

// one thread per element, i fastest (coalesced stores of A and loads of Wtrue)
// general block: rows [row0, row0+nr) of the n_global-row matrix; Wt holds the nr local rows (leading dimension nr)
__global__ void __launch_bounds__(256)
k_synth_block(double* __restrict__ A, const double* __restrict__ Wt, const double* __restrict__ Ht, int64_t n_global,
              int64_t row0, int64_t nr, int k, int64_t col0, uint64_t base, double noise, double na_frac)
{
    extern __shared__ double hs[];
    const int64_t j = blockIdx.y;
    for (int c = threadIdx.x; c < k; c += blockDim.x) hs[c] = Ht[c + (int64_t)k * j];
    __syncthreads();
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < k; c++) s = fma(Wt[i + nr * c], hs[c], s);
        const uint64_t idx = (uint64_t)(row0 + i) + (uint64_t)n_global * (uint64_t)(col0 + j);
        s = fma(noise, splitmix_u(base + 3, idx), s);
        if (na_frac > 0.0 && splitmix_u(base + 4, idx) < na_frac) s = nanv;
        A[i + nr * j] = s;
    }
}

__global__ void k_uniform_rows(double* __restrict__ out, int64_t n_global, int64_t row0, int64_t nr, int k, uint64_t seed)
{
    const int64_t total = nr * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % nr, c = e / nr;
        out[e] = splitmix_u(seed, (uint64_t)(row0 + i) + (uint64_t)n_global * (uint64_t)c);
    }
}

__global__ void __launch_bounds__(256)
k_synth(double* __restrict__ A, const double* __restrict__ Wt, const double* __restrict__ Ht, int64_t n, int64_t m,
        int k, int64_t col0, uint64_t base, double noise, double na_frac)
{
    extern __shared__ double hs[];      // Htrue column of this block's j
    const int64_t j = blockIdx.y;
    for (int c = threadIdx.x; c < k; c += blockDim.x) hs[c] = Ht[c + (int64_t)k * j];
    __syncthreads();
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < k; c++) s = fma(Wt[i + n * c], hs[c], s);
        const uint64_t idx = (uint64_t)i + (uint64_t)n * (uint64_t)(col0 + j);
        s = fma(noise, splitmix_u(base + 3, idx), s);
        if (na_frac > 0.0 && splitmix_u(base + 4, idx) < na_frac) s = nanv;
        A[i + n * j] = s;
    }
}

}  // namespace

void launch_uniform(double* out, int64_t count, uint64_t seed, uint64_t offset, double scale, cudaStream_t st)
{
    if (count <= 0) return;
    const int grid = (int)std::min<int64_t>(ceil_div(count, 256), 148 * 8);
    k_uniform<<<grid, 256, 0, st>>>(out, count, seed, offset, scale);
    NNLM_LAUNCHED();
}

void launch_synth(double* A, int64_t n, int64_t m, int k, int64_t col0, uint64_t base, double noise, double na_frac,
                  cudaStream_t st)
{
    NNLM_REQUIRE(m <= 65535 * 64ll, "synthetic generator: too many columns per call");
    DevBuf<double> Wt((size_t)n * k), Ht((size_t)k * m);
    launch_uniform(Wt.p, n * k, base + 1, 0, 1.0, st);
    launch_uniform(Ht.p, (int64_t)k * m, base + 2, (uint64_t)k * (uint64_t)col0, 1.0, st);
    // grid.y is limited to 65535: process column bands
    for (int64_t j0 = 0; j0 < m; j0 += 65535) {
        const int64_t mc = std::min<int64_t>(65535, m - j0);
        dim3 grid((unsigned)std::min<int64_t>(ceil_div(n, 256), 64), (unsigned)mc);
        k_synth<<<grid, 256, sizeof(double) * k, st>>>(A + n * j0, Wt.p, Ht.p + (int64_t)k * j0, n, mc, k, col0 + j0, base,
                                                      noise, na_frac);
        NNLM_LAUNCHED();
    }
    NNLM_CUDA_CHECK(cudaStreamSynchronize(st));    // Wt/Ht are freed on return
}

void launch_synth_block(double* A, int64_t n_global, int64_t row0, int64_t nr, int64_t col0, int64_t mc, int k, uint64_t base,
                        double noise, double na_frac, cudaStream_t st)
{
    if (nr <= 0 || mc <= 0) return;
    DevBuf<double> Wt((size_t)nr * k), Ht((size_t)k * mc);
    k_uniform_rows<<<(int)std::min<int64_t>(ceil_div(nr * k, 256), 148 * 8), 256, 0, st>>>(Wt.p, n_global, row0, nr, k, base + 1);
    NNLM_LAUNCHED();
    launch_uniform(Ht.p, (int64_t)k * mc, base + 2, (uint64_t)k * (uint64_t)col0, 1.0, st);
    for (int64_t j0 = 0; j0 < mc; j0 += 65535) {
        const int64_t cnt = std::min<int64_t>(65535, mc - j0);
        dim3 grid((unsigned)std::min<int64_t>(ceil_div(nr, 256), 64), (unsigned)cnt);
        k_synth_block<<<grid, 256, sizeof(double) * k, st>>>(A + nr * j0, Wt.p, Ht.p + (int64_t)k * j0, n_global, row0, nr, k,
                                                            col0 + j0, base, noise, na_frac);
        NNLM_LAUNCHED();
    }
    NNLM_CUDA_CHECK(cudaStreamSynchronize(st));
}

}  // namespace nnlm
