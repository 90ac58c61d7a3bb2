// pair_bench.cu — scd_pair.cuh vs scd_chain.cuh on a synthetic NNLS batch (k = 50): time per launch, bitwise comparison, and the
// phase clocks of one pair (NNLM_PAIR_PROF).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I nnlm_b200/csrc -I scratch -I scratch
#define NNLM_TEAM_PROF 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "scd_team.cuh"
using namespace nnlm;
#ifndef KK
#define KK 50
#endif
constexpr int NHH = (KK + 3) / 4;
int main(int argc, char** argv)
{
    const int k = KK;
    std::vector<int64_t> sizes = {1184, 10000, 50000};
    if (argc > 1) { sizes.clear(); for (int i = 1; i < argc; i++) sizes.push_back(atoll(argv[i])); }
    std::mt19937_64 rng(5);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const int nf = 400;
    std::vector<double> F((size_t)k * nf), G((size_t)k * k, 0.0);
    for (auto& x : F) x = U(rng);
    for (int a = 0; a < k; a++) for (int b = 0; b < k; b++) { double s = 0; for (int i = 0; i < nf; i++) s += F[a * nf + i] * F[b * nf + i]; G[a + k * b] = s + (a == b ? 1e-16 : 0.0); }
    const int64_t maxc = 50000;
    std::vector<double> Q((size_t)maxc * k), X0((size_t)maxc * k);
    for (int64_t j = 0; j < maxc; j++) {
        double xt[128];
        for (int a = 0; a < k; a++) xt[a] = U(rng) < 0.5 ? 0.0 : U(rng);
        for (int a = 0; a < k; a++) { double s = 0; for (int b = 0; b < k; b++) s += G[a + k * b] * xt[b]; Q[j * k + a] = s + 2.0 * (U(rng) - 0.5); }
        for (int a = 0; a < k; a++) X0[j * k + a] = 0.01 * U(rng);
    }
    double *dG, *dQ, *dX, *dX0, *dY; unsigned long long* dS; unsigned int* dC;
    cudaMalloc(&dG, G.size() * 8); cudaMalloc(&dQ, Q.size() * 8); cudaMalloc(&dX, X0.size() * 8); cudaMalloc(&dX0, X0.size() * 8); cudaMalloc(&dY, X0.size() * 8);
    cudaMalloc(&dS, 8); cudaMalloc(&dC, 16);
    cudaMemcpy(dG, G.data(), G.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dQ, Q.data(), Q.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dX0, X0.data(), X0.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<double> ref(X0.size()), got(X0.size());
    for (int64_t nc : sizes) {
        for (int variant = 0; variant < 7; variant++) {
            float best = 1e9; unsigned long long sw = 0;
            for (int rep = 0; rep < 4; rep++) {
                cudaMemcpy(dX, dX0, nc * k * 8, cudaMemcpyDeviceToDevice); cudaMemset(dS, 0, 8);
                cudaEventRecord(e0);
                try {
                    if (variant == 0) scd_chain::launch<NHH, 1>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                    if (variant == 1) scd_chain::launch<NHH, 2>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                    if (variant == 2) scd_team::launch<NHH, 1, 1>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                    if (variant == 3) scd_team::launch<NHH, 2, 1>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                    if (variant == 4) scd_team::launch<NHH, 2, 2>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                    if (variant == 5) scd_team::launch<NHH, 4, 2>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                    if (variant == 6) scd_team::launch<NHH, 4, 4>(dX, dG, dQ, 1, nullptr, k, nc, 0.0, 50, 1e-9, dS, dC, 0);
                } catch (const std::exception& e) { printf("variant %d: %s\n", variant, e.what()); break; }
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                if (err != cudaSuccess) { printf("variant %d: sync error %s\n", variant, cudaGetErrorString(err)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
                cudaMemcpy(&sw, dS, 8, cudaMemcpyDeviceToHost);
            }
            cudaMemcpy(got.data(), dX, nc * k * 8, cudaMemcpyDeviceToHost);
            if (variant == 0) ref = got;
            size_t diff = 0; for (int64_t e = 0; e < nc * k; e++) diff += memcmp(&ref[e], &got[e], 8) != 0;
            long long prof[16] = {0};
            if (variant >= 2) cudaMemcpyFromSymbol(prof, scd_team::g_team_prof, sizeof prof);
            const char* names[7] = {"chain<1>  ", "chain<2>  ", "team<1,1> ", "team<2,1> ", "team<2,2> ", "team<4,2> ", "team<4,4> "};
            printf("ncol %6lld %s %8.4f ms  sweeps/col %.2f  differing %zu", (long long)nc, names[variant], best, (double)sw / nc, diff);
            if (variant >= 2) {
                const double nblk = (double)prof[8] * ((NHH + 1) / 2);
                printf("  | per block: mma0 deferred+wait %.0f crit %.0f; mma1 %.0f %.0f | chain init %.0f steps %.0f publish %.0f pre %.0f",
                       prof[0] / nblk, prof[1] / nblk, prof[2] / nblk, prof[3] / nblk, prof[9] / nblk, prof[10] / nblk, prof[11] / nblk, prof[12] / nblk);
            }
            printf("\n");
        }
    }
    return 0;
}
