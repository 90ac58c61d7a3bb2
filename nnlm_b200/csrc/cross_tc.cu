// cross_tc.cu — K2 on the 5th-generation tensor cores: the cross-product  Q = Wt * A  (k x ncol) of one half-iteration,
// which the reference forms one column at a time as `Wt * A.col(j)` (src/update_with_missing.cpp:39,45), as ONE
// HBM-bound pass over the resident copy of A.
//
// Precision. The reference is fp64; the parity bound is 1e-5 on W and H. A and the fixed factor are each stored as two
// fp16 planes, x*s = hi + lo*2^-11 (s a power of two, 22-24 significant bits, the same 4 bytes per element of A as fp32):
//     sum_i A[i,j] F[a,i]  =  ( sum hi_A hi_F  +  2^-11 * sum (hi_A lo_F + lo_A hi_F) ) / (s_A s_F[a])      (+ O(2^-22) dropped)
// Each product of two fp16 values is exact in the fp32 accumulator; the two sums live in separate TMEM accumulators so
// the small terms are not rounded against the large ones. Long fp32 accumulation is avoided (the tensor core truncates
// each accumulate, which biases long sums): every `drain` k-blocks (4 = 256 contraction indices; 2 for the 128-row factor tile) the accumulators are read out of TMEM and added into fp64 registers, while the tensor core
// continues into the other TMEM buffer. Measured error of the whole cross-product vs fp64: see tests/test_gpu_cross.py.
//
// Centering. The tensor core truncates every accumulate; with all-positive data that is a systematic bias (measured:
// -1.2e-6 relative at 256 indices per drain, -1.6e-5 at 4096). The planes therefore hold A minus its column mean,
// A'[i,j] = A[i,j] - c[j], whose products have mixed signs (accumulators stay small, truncation is unbiased), and the
// mean component is added back exactly in fp64:  sum_i A[i,j] F[a,i] = sum_i A'[i,j] F[a,i] + c[j] * rowsum(F)[a].
//
// Mapping (UMMA D[M x N] = A[M x K] * B[N x K]^T, both operands K-major, SWIZZLE_128B):
//   M = 128 columns j of A   (operand rows = columns of the matrix, contiguous along the contraction index i)
//   N = NP  = k padded to 32/64/128 (rows a of the factor, row-major copy made by split_factor)
//   K = 64 fp16 (= one 128-byte swizzle row) per pipeline stage, UMMA_K = 16 -> 4 k-steps x 3 products per stage
// Work = (tile, k-block) units in tile-major order, cut into gridDim.x equal contiguous ranges (stream-K): every CTA
// streams the same number of bytes; a tile shared by several CTAs receives one fp64 partial per CTA in slot order
// (deterministic), summed by the solver exactly like the split-K partials of the fp64 path.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..9 = epilogue (TMEM -> fp64).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace nnlm {

using namespace tc;

namespace {

constexpr int BM = 128;          // columns of A per tile
constexpr int BK = 64;           // fp16 elements per k-block (128 bytes)
// k-blocks accumulated in TMEM between two drains into the fp64 registers. The fp32 accumulation in TMEM is the dominant error
// of the whole cross-product (measured, 5000 x 2000, contraction length 2000: relative Frobenius error 6.4e-10 / 3.5e-10 /
// 2.0e-10 / 1.5e-10 at 16 / 4 / 2 / 1 k-blocks per drain); the epilogue hides behind the HBM stream down to 4 k-blocks for the
// 32- and 64-row factor tiles (cross-product time +0.9 %); the 128-row tile drains every k-block, which buys the 1e-5 parity of
// the first iteration at k = 128 (the near-rank-one tiny init amplifies the W-half's error ~2000x into H:
// tests/test_gpu_scale_parity.py, profiles/r2_b_drain_interval.md). Only the hi*hi accumulator is drained that often; the
// small-term accumulator (2^-11 of it) keeps 16 k-blocks per drain, which halves the epilogue's conversions.
constexpr int DRAIN_SMALL = 4;    // NP = 32, 64: 256 contraction indices
constexpr int DRAIN_LARGE = 1;    // NP = 128:     64 contraction indices
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 32 * (2 + EPI_WARPS);

struct CrossParams {
    int64_t ncol;          // columns of A (rows of the UMMA M operand)
    int64_t kblocks;       // ceil(len / BK)
    int64_t units;         // tiles * kblocks
    int k;                 // true rank
    int slots;             // slots of Qp (most CTAs that touch one tile)
    double* Qp;            // [slots][ncol][k]
    const double* unscale; // [NP]: 1 / (s_A * s_F[a])
    int drain;             // k-blocks accumulated in TMEM between two drains into the fp64 registers
    int d1_every;          // drains of the large-term accumulator per drain of the small-term one (MODE 0; 1 otherwise)
    int halves;            // 2: every k-block is drained twice, after 32 indices each (needs drain == 1); else 1
    const double* center;  // [ncol] mean of each column of A that was subtracted before the split (or nullptr)
    const double* fsum;    // [k] row sums of the factor
};

// first CTA whose unit range [floor(c*U/P), floor((c+1)*U/P)) reaches unit `u`
__host__ __device__ inline int64_t first_cta_of_unit(int64_t u, int64_t U, int64_t P) {
    int64_t c = (u * P) / U;
    while (c + 1 < P && ((c + 1) * U) / P <= u) c++;
    while (c > 0 && (c * U) / P > u) c--;
    return c;
}

// MODE 0: the three-product hi/lo scheme above. MODE 1: ONE product per k-step from the `hi` planes only, for operands that
// are small integers stored in fp16 (the 0/1 missing mask and 11-bit fixed-point slices of the NA path, na_gram.cu): every
// product and every fp32 partial sum below 2^24 is exact, so the result is exact up to the slicing. MODE 2: the same with
// TWO factor planes per pass (two consecutive slices, 2^11 apart: f_hi, f_lo) against the one `hi` plane of A, combined
// like MODE 0 in the epilogue — the mask plane is streamed once for two slices.
template <int NP, int STAGES, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
k_cross_tc(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
           const __grid_constant__ CUtensorMap mapF_hi, const __grid_constant__ CUtensorMap mapF_lo, const CrossParams p)
{
    constexpr int A_BYTES = BM * BK * 2;                   // 16 KB per plane per stage
    constexpr int F_BYTES = NP * BK * 2;
    constexpr int STAGE_BYTES = MODE == 0 ? 2 * A_BYTES + 2 * F_BYTES : (MODE == 1 ? A_BYTES + F_BYTES : A_BYTES + 2 * F_BYTES);
    constexpr int CPT = NP / 2;                            // accumulator columns per epilogue thread
    constexpr uint32_t TMEM_COLS = (4 * NP <= 128) ? 128 : (4 * NP <= 256 ? 256 : 512);
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // f16 x f16 -> f32, K-major

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment: align by hand (the launch reserves the slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* tiles = smem;                                                   // STAGES * STAGE_BYTES
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t P = gridDim.x, U = p.units, KBn = p.kblocks;
    const int64_t u0 = ((int64_t)blockIdx.x * U) / P, u1 = ((int64_t)(blockIdx.x + 1) * U) / P;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], EPI_WARPS); }
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
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t u = u0; u < u1; u++) {
                const int64_t tile = u / KBn, kb = u % KBn;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = tiles + stage * STAGE_BYTES;
                mbar_expect_tx(&full[stage], STAGE_BYTES);
                const int c0 = (int)(kb * BK), c1 = (int)(tile * BM);
                if (MODE == 0) {
                    tma_load_2d(sa, &mapA_hi, &full[stage], c0, c1);
                    tma_load_2d(sa + A_BYTES, &mapA_lo, &full[stage], c0, c1);
                    tma_load_2d(sa + 2 * A_BYTES, &mapF_hi, &full[stage], c0, 0);
                    tma_load_2d(sa + 2 * A_BYTES + F_BYTES, &mapF_lo, &full[stage], c0, 0);
                } else {
                    tma_load_2d(sa, &mapA_hi, &full[stage], c0, c1);
                    tma_load_2d(sa + A_BYTES, &mapF_hi, &full[stage], c0, 0);
                    if (MODE == 2) tma_load_2d(sa + A_BYTES + F_BYTES, &mapF_lo, &full[stage], c0, 0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t chunk = 0;                       // running drain-chunk counter of this CTA
            uint32_t big = 0, cur_big = 0;            // MODE 0: the small-term accumulator d1 is drained once per p.d1_every chunks
            int64_t u = u0;
            while (u < u1) {
                const int64_t tile = u / KBn;
                const int64_t seg_end = min(u1, (tile + 1) * KBn);
                uint32_t sub = 0;
                while (p.halves == 2 && u < seg_end) {            // 32 contraction indices per fp32 accumulation: two chunks per k-block
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
                    const uint64_t f_hi = make_desc(sa + (MODE == 0 ? 2 * A_BYTES : A_BYTES));
                    const uint64_t f_lo = make_desc(sa + (MODE == 0 ? 2 * A_BYTES + F_BYTES : A_BYTES + F_BYTES));
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const uint32_t buf = chunk & 1, tph = (chunk >> 1) & 1;
                        mbar_wait(&tempty[buf], tph ^ 1);
                        tc_fence_after();
                        const bool big_first = (sub % (uint32_t)p.d1_every) == 0;
                        if (big_first) cur_big = big++;
                        const uint32_t d0 = tmem_base + buf * NP, d1 = tmem_base + 2 * NP + (cur_big & 1) * NP;
#pragma unroll
                        for (int ks = 2 * hh; ks < 2 * hh + 2; ks++) {
                            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                            const uint32_t acc = (ks == 2 * hh) ? 0u : 1u;
                            const uint32_t acc1 = (ks == 2 * hh && big_first) ? 0u : 1u;
                            umma_f16(d0, a_hi + adv, f_hi + adv, IDESC, acc);
                            if (MODE != 1) umma_f16(d1, a_hi + adv, f_lo + adv, IDESC, acc1);
                            if (MODE == 0) umma_f16(d1, a_lo + adv, f_hi + adv, IDESC, 1u);
                        }
                        if (hh == 1) tc_commit(&empty[stage]);
                        tc_commit(&tfull[buf]);
                        chunk++;
                        sub++;
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    u++;
                }
                while (u < seg_end) {
                    const int64_t chunk_end = min(seg_end, u + (int64_t)p.drain);
                    const uint32_t buf = chunk & 1, tph = (chunk >> 1) & 1;
                    mbar_wait(&tempty[buf], tph ^ 1);           // epilogue has drained this TMEM buffer
                    tc_fence_after();
                    const bool big_first = (sub % (uint32_t)p.d1_every) == 0;
                    if (big_first) cur_big = big++;
                    // TMEM: d0[2] at columns [0, 2 NP), d1[2] at [2 NP, 4 NP). The d1 buffer of big chunk B is free again when B+2
                    // starts: the tempty wait above covers the drain chunk two before this one, which is at or after the last
                    // chunk of B (B+1 holds at least one chunk), and the epilogue reads d1 before it arrives on that chunk.
                    const uint32_t d0 = tmem_base + buf * NP, d1 = tmem_base + 2 * NP + (cur_big & 1) * NP;
                    bool first = true;
                    for (; u < chunk_end; u++) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
                        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
                        const uint64_t f_hi = make_desc(sa + (MODE == 0 ? 2 * A_BYTES : A_BYTES));
                        const uint64_t f_lo = make_desc(sa + (MODE == 0 ? 2 * A_BYTES + F_BYTES : A_BYTES + F_BYTES));
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ks++) {
                            const uint64_t adv = (uint64_t)((ks * 32) >> 4);     // 16 fp16 = 32 bytes along K
                            const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                            const uint32_t acc1 = (first && ks == 0 && big_first) ? 0u : 1u;
                            umma_f16(d0, a_hi + adv, f_hi + adv, IDESC, acc);    // hi*hi
                            if (MODE != 1) umma_f16(d1, a_hi + adv, f_lo + adv, IDESC, acc1);   // hi*lo (MODE 2: the next slice)
                            if (MODE == 0) umma_f16(d1, a_lo + adv, f_hi + adv, IDESC, 1u);      // lo*hi
                        }
                        first = false;
                        tc_commit(&empty[stage]);               // frees the smem stage when these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&tfull[buf]);                     // accumulators of this chunk are complete
                    chunk++;
                    sub++;
                }
            }
        }
    } else {
        // ===================================== epilogue: TMEM -> fp64 registers -> partial tile =====================================
        const int ew = warp - 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = ew >> 2;                     // which half of the NP accumulator columns
        const int row = quarter * 32 + lane;          // row of the tile = column of A
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t chunk = 0, big = 0, cur_big = 0;
        int64_t u = u0;
        while (u < u1) {
            const int64_t tile = u / KBn;
            const int64_t seg_end = min(u1, (tile + 1) * KBn);
            double acc[CPT];
#pragma unroll
            for (int c = 0; c < CPT; c++) acc[c] = 0.0;
            uint32_t sub = 0;
            int hh = 0;                                 // p.halves == 2: which half of the k-block this chunk is
            while (u < seg_end) {
                const bool half_mode = p.halves == 2;
                const int64_t chunk_end = half_mode ? u + 1 : min(seg_end, u + (int64_t)p.drain);
                const bool closes_kblock = !half_mode || hh == 1;
                const uint32_t buf = chunk & 1, tph = (chunk >> 1) & 1;
                if ((sub % (uint32_t)p.d1_every) == 0) cur_big = big++;
                // the small-term accumulator is read with the last drain chunk of its group (or of the tile)
                const bool with_d1 = MODE != 1 && ((closes_kblock && chunk_end == seg_end) || ((sub + 1) % (uint32_t)p.d1_every) == 0);
                mbar_wait(&tfull[buf], tph);
                tc_fence_after();
                constexpr int CH = CPT > 32 ? 32 : CPT;          // TMEM columns read per tcgen05.ld
                const uint32_t t0 = tmem_base + lane_addr + buf * NP + half * CPT;
                const uint32_t t1 = tmem_base + lane_addr + 2 * NP + (cur_big & 1) * NP + half * CPT;
                // one read-out array live at a time (the 128-row tile has 64 fp64 accumulators per thread already)
#pragma unroll
                for (int ch = 0; ch < CPT / CH; ch++) {
                    uint32_t r[CH];
                    TmemLd<CH>::ld(t0 + ch * CH, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < CH; c++) acc[ch * CH + c] += (double)__uint_as_float(r[c]);
                }
                if (with_d1) {
#pragma unroll
                    for (int ch = 0; ch < CPT / CH; ch++) {
                        uint32_t r[CH];
                        TmemLd<CH>::ld(t1 + ch * CH, r);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < CH; c++) acc[ch * CH + c] = fma((double)__uint_as_float(r[c]), LO_UNSCALE, acc[ch * CH + c]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);           // buffer may be overwritten by the next-but-one chunk
                if (closes_kblock) { u = chunk_end; hh = 0; } else hh = 1;
                chunk++;
                sub++;
            }
            // partial tile -> slot (index of this CTA among the CTAs that touch the tile)
            const int64_t slot = (int64_t)blockIdx.x - first_cta_of_unit(tile * KBn, U, P);
            const int64_t j = tile * BM + row;
            if (j < p.ncol) {
                double* out = p.Qp + (slot * p.ncol + j) * p.k;
                const double cj = (slot == 0 && p.center != nullptr) ? p.center[j] : 0.0;    // mean component, added once
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    const int a = half * CPT + c;
                    if (a < p.k) out[a] = MODE == 0 ? fma(cj, p.fsum[a], acc[c] * p.unscale[a]) : acc[c] * p.unscale[a];
                }
                // the CTA that finishes a tile zero-fills the slots nobody writes (consumers sum all p.slots of them): no
                // memset of the whole partial buffer per launch
                if (seg_end == (tile + 1) * KBn) {
                    for (int64_t sp = slot + 1; sp < p.slots; sp++) {
                        double* z = p.Qp + (sp * p.ncol + j) * p.k;
#pragma unroll
                        for (int c = 0; c < CPT; c++) {
                            const int a = half * CPT + c;
                            if (a < p.k) z[a] = 0.0;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}



// ---------------------------------------------------------------------------------------------------- pair contraction
// k_mask_tc2: the exact integer contraction of the NA path (MODE 2's arithmetic) by a PAIR of CTAs (tcgen05 cta_group::2).
// Why: k_cross_tc<128,4,2> moves 48 KB through L2 per k-block for 4.2 Mflop (87 flop/B) and sits on the chip's L2 -> SM ingest
// cap (6200 of ~6300 B/clk: profiles/r2_d_config4.md). A pair forms D[256 x 256] = M[256 x K] * Z[256 x K]^T per k-step: CTA r
// holds mask columns tile*256 + r*128 .. +128 (its half of the M operand and of the accumulator rows) and loads slice tile r
// (its half of the N operand: N = slice 0 | slice 1); the tensor cores of both SMs read both halves of N. Per CTA and k-block
// 32 KB for 4.2 Mflop (131 flop/B); TMEM per CTA 256 columns x 2 buffers; epilogue and fp64 register accumulators exactly as
// MODE 2. The leader CTA (rank 0) issues the MMAs and owns the `full` barriers (both CTAs' TMA loads complete on them);
// `empty` / `tfull` exist in both CTAs and are signalled by multicast commits; the peer's epilogue arrives remotely on the
// leader's `tempty`.
constexpr int P2_STAGES = 6;

__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a pair: the bytes complete on the barrier at cluster address `bar_cluster` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {       // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_mask_tc2(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapF0,
           const __grid_constant__ CUtensorMap mapF1, const CrossParams p)
{
    constexpr int NP = 128;                                // Z columns per slice tile
    constexpr int T_BYTES = 128 * BK * 2;                  // 16 KB: one 128-row operand tile per stage
    constexpr int STAGE_BYTES = 2 * T_BYTES;               // this CTA's half of M and its half of N
    constexpr int CPT = NP / 2;
    // f16 x f16 -> f32, K-major, N = 256, M = 256
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + P2_STAGES * STAGE_BYTES);
    uint64_t* empty = full + P2_STAGES;
    uint64_t* tfull = empty + P2_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int64_t P = gridDim.x / 2, pair = blockIdx.x / 2, U = p.units, KBn = p.kblocks;
    const int64_t u0 = (pair * U) / P, u1 = ((pair + 1) * U) / P;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P2_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const CUtensorMap* mapF = rank == 0 ? &mapF0 : &mapF1;
            for (int64_t u = u0; u < u1; u++) {
                const int64_t tile = u / KBn, kb = u % KBn;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = tiles + stage * STAGE_BYTES;
                const uint32_t bar0 = mapa_rank(smem_u32(&full[stage]), 0);
                if (rank == 0) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
                const int c0 = (int)(kb * BK), c1 = (int)(tile * 256 + rank * 128);
                tma_load_2d_pair(sa, &mapA, bar0, c0, c1);                         // (rows past the end read as zero)
                tma_load_2d_pair(sa + T_BYTES, mapF, bar0, c0, 0);
                if (++stage == P2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================================
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t chunk = 0;
            int64_t u = u0;
            while (u < u1) {
                const int64_t tile = u / KBn;
                const int64_t seg_end = min(u1, (tile + 1) * KBn);
                while (u < seg_end) {
                    const int64_t chunk_end = min(seg_end, u + (int64_t)p.drain);
                    const uint32_t buf = chunk & 1, tph = (chunk >> 1) & 1;
                    mbar_wait(&tempty[buf], tph ^ 1);           // both epilogues have drained this TMEM buffer
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * 256;
                    bool first = true;
                    for (; u < chunk_end; u++) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
                        const uint64_t a = make_desc(sa), f = make_desc(sa + T_BYTES);
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ks++) {
                            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                            umma_f16_pair(d, a + adv, f + adv, IDESC, (first && ks == 0) ? 0u : 1u);
                        }
                        first = false;
                        tc_commit_pair(&empty[stage]);          // frees the stage in both CTAs when these MMAs retire
                        if (++stage == P2_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit_pair(&tfull[buf]);                // accumulators of this chunk are complete (both CTAs)
                    chunk++;
                }
            }
        }
    } else {
        // ===================================== epilogue (both CTAs): TMEM -> fp64 registers -> partial tile =====================================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t chunk = 0;
        int64_t u = u0;
        while (u < u1) {
            const int64_t tile = u / KBn;
            const int64_t seg_end = min(u1, (tile + 1) * KBn);
            double acc[CPT];
#pragma unroll
            for (int c = 0; c < CPT; c++) acc[c] = 0.0;
            while (u < seg_end) {
                const int64_t chunk_end = min(seg_end, u + (int64_t)p.drain);
                const uint32_t buf = chunk & 1, tph = (chunk >> 1) & 1;
                mbar_wait(&tfull[buf], tph);
                tc_fence_after();
                const uint32_t t0 = tmem_base + lane_addr + buf * 256 + half * CPT;        // slice 0 sums
                const uint32_t t1 = t0 + NP;                                               // slice 1 sums
#pragma unroll
                for (int ch = 0; ch < CPT / 32; ch++) {
                    uint32_t r[32];
                    TmemLd<32>::ld(t0 + ch * 32, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 32; c++) acc[ch * 32 + c] += (double)__uint_as_float(r[c]);
                }
#pragma unroll
                for (int ch = 0; ch < CPT / 32; ch++) {
                    uint32_t r[32];
                    TmemLd<32>::ld(t1 + ch * 32, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 32; c++) acc[ch * 32 + c] = fma((double)__uint_as_float(r[c]), LO_UNSCALE, acc[ch * 32 + c]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[buf]), 0));   // on the leader's barrier
                u = chunk_end;
                chunk++;
            }
            const int64_t slot = pair - first_cta_of_unit(tile * KBn, U, P);
            const int64_t j = tile * 256 + (int64_t)rank * 128 + row;
            if (j < p.ncol) {
                double* out = p.Qp + (slot * p.ncol + j) * p.k;
#pragma unroll
                for (int c = 0; c < CPT; c++) out[half * CPT + c] = acc[c] * p.unscale[half * CPT + c];
                if (seg_end == (tile + 1) * KBn) {            // the pair that finishes a tile zero-fills the slots nobody writes
                    for (int64_t sp = slot + 1; sp < p.slots; sp++) {
                        double* z = p.Qp + (sp * p.ncol + j) * p.k;
#pragma unroll
                        for (int c = 0; c < CPT; c++) z[half * CPT + c] = 0.0;
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                       // nobody leaves while the partner may still read its shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------- operand preparation
// |x| maxima as fp64 bit patterns (non-negative doubles order like unsigned integers)
__global__ void k_rowmax(const double* __restrict__ F, int k, int64_t len, unsigned long long* __restrict__ rowmax)
{
    extern __shared__ unsigned long long smx[];       // [k]
    for (int a = threadIdx.x; a < k; a += blockDim.x) smx[a] = 0ull;
    __syncthreads();
    const int64_t total = (int64_t)k * len;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs(F[e]);
        if (!is_missing(v)) atomicMax(&smx[e % k], (unsigned long long)__double_as_longlong(v));
    }
    __syncthreads();
    for (int a = threadIdx.x; a < k; a += blockDim.x) if (smx[a]) atomicMax(&rowmax[a], smx[a]);
}

// scales[a] = s_F[a]; unscale[a] = 1 / (s_A * s_F[a])
__global__ void k_make_scales(const unsigned long long* __restrict__ rowmax, int k, int np, const double* __restrict__ sA,
                              double* __restrict__ scales, double* __restrict__ unscale)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= np) return;
    const double s = (a < k) ? pow2_scale(__longlong_as_double((long long)rowmax[a])) : 1.0;
    scales[a] = s;
    unscale[a] = 1.0 / (sA[0] * s);
}

// F (k x len, column-major fp64) -> planes [NP][ld] (row a contiguous along i), rows >= k zero
__global__ void __launch_bounds__(256)
k_split_factor(const double* __restrict__ F, int k, int64_t len, int64_t ld, int np, const double* __restrict__ scales,
               __half* __restrict__ hi, __half* __restrict__ lo)
{
    extern __shared__ double tile[];                   // [64][k + 1]
    const int64_t i0 = (int64_t)blockIdx.x * 64;
    const int cnt = (int)min((int64_t)64, len - i0);
    const int kp = k + 1;
    for (int e = threadIdx.x; e < cnt * k; e += 256) tile[(e / k) * kp + (e % k)] = F[(int64_t)k * i0 + e];
    __syncthreads();
    for (int e = threadIdx.x; e < np * 64; e += 256) {
        const int a = e / 64, ii = e % 64;
        if (i0 + ii >= ld) continue;
        __half h = __float2half_rn(0.f), l = h;
        if (a < k && ii < cnt) split2(tile[ii * kp + a] * scales[a], h, l);
        hi[(int64_t)a * ld + i0 + ii] = h;
        lo[(int64_t)a * ld + i0 + ii] = l;
    }
}

// max |A| over finite entries -> bit pattern (one atomicMax per block)
__global__ void k_absmax(const double* __restrict__ A, int64_t total, unsigned long long* __restrict__ out)
{
    __shared__ unsigned long long red[32];
    unsigned long long m = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs(A[e]);
        if (!is_missing(v)) m = max(m, (unsigned long long)__double_as_longlong(v));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = max(m, red[w]);
        if (m) atomicMax(out, m);
    }
}

// centred values are bounded by 2 max|A|
__global__ void k_scale_from_max(const unsigned long long* __restrict__ mx, double* __restrict__ sA)
{
    sA[0] = pow2_scale(2.0 * __longlong_as_double((long long)mx[0]));
}

// mean over the finite entries of every column (warp per column) and of every row (thread per row)
__global__ void k_col_means(const double* __restrict__ A, int64_t len, int64_t ncol, double* __restrict__ mean)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < ncol; j += nwarp) {
        double s = 0.0, c = 0.0;
        for (int64_t i = lane; i < len; i += 32) {
            const double a = A[i + len * j];
            if (!is_missing(a)) { s += a; c += 1.0; }
        }
        s = warp_sum(s); c = warp_sum(c);
        if (lane == 0) mean[j] = c > 0 ? s / c : 0.0;
    }
}

__global__ void k_row_means(const double* __restrict__ A, int64_t len, int64_t ncol, double* __restrict__ mean)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s = 0.0, c = 0.0;
    for (int64_t j = 0; j < ncol; j++) {
        const double a = A[i + len * j];
        if (!is_missing(a)) { s += a; c += 1.0; }
    }
    mean[i] = c > 0 ? s / c : 0.0;
}

// A (len x ncol, column-major fp64) -> planes of A - colmean (pitch ld_a, row j) and of A' - rowmean (pitch ld_t, row i);
// non-finite entries become 0 in the planes (= the mean)
__global__ void __launch_bounds__(256)
k_split_matrix(const double* __restrict__ A, int64_t len, int64_t ncol, const double* __restrict__ sA,
               const double* __restrict__ colmean, const double* __restrict__ rowmean,
               __half* __restrict__ a_hi, __half* __restrict__ a_lo, int64_t ld_a,
               __half* __restrict__ t_hi, __half* __restrict__ t_lo, int64_t ld_t)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    const double s = sA[0];
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = i0 + tx, j = j0 + r;
        double a = nanv;
        if (i < len && j < ncol) {
            a = A[i + len * j];
            if (a_hi) {
                __half h = __float2half_rn(0.f), l = h;
                if (!is_missing(a)) split2((a - colmean[j]) * s, h, l);
                a_hi[i + ld_a * j] = h;
                a_lo[i + ld_a * j] = l;
            }
        }
        tile[r][tx] = a;
    }
    __syncthreads();
    if (t_hi) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t j = j0 + tx, i = i0 + r;
            if (i < len && j < ncol) {
                const double a = tile[tx][r];
                __half h = __float2half_rn(0.f), l = h;
                if (!is_missing(a)) split2((a - rowmean[i]) * s, h, l);
                t_hi[j + ld_t * i] = h; t_lo[j + ld_t * i] = l;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host side
template <int NP, int STAGES, int MODE>
void launch_np(const CrossPlan& plan, const __half* a_hi, const __half* a_lo, const __half* f_hi, const __half* f_lo,
               const double* unscale, const double* center, const double* fsum, double* Qp, int drain, cudaStream_t st)
{
    constexpr size_t stage = MODE == 0 ? (2 * BM * BK * 2 + 2 * NP * BK * 2) : (MODE == 1 ? (BM * BK * 2 + NP * BK * 2) : (BM * BK * 2 + 2 * NP * BK * 2));
    constexpr size_t smem = (size_t)STAGES * stage + 1024 /*alignment slack*/ + 256;
    auto kern = k_cross_tc<NP, STAGES, MODE>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const CUtensorMap mA_hi = make_map(a_hi, plan.len, plan.ncol, plan.ld_a, BM);
    const CUtensorMap mA_lo = MODE == 0 ? make_map(a_lo, plan.len, plan.ncol, plan.ld_a, BM) : mA_hi;
    const CUtensorMap mF_hi = make_map(f_hi, plan.len, NP, plan.ld_f, NP);
    const CUtensorMap mF_lo = MODE != 1 ? make_map(f_lo, plan.len, NP, plan.ld_f, NP) : mF_hi;
    CrossParams p;
    p.ncol = plan.ncol; p.kblocks = plan.kblocks; p.units = plan.units; p.k = plan.k; p.slots = plan.slots; p.Qp = Qp; p.unscale = unscale;
    static const int drain_env = [] { const char* e = getenv("NNLM_TC_DRAIN"); return e ? atoi(e) : 0; }();
    p.drain = drain > 0 ? drain : (drain_env > 0 ? drain_env : (NP >= 128 ? DRAIN_LARGE : DRAIN_SMALL));
    // the hi*lo + lo*hi accumulator is 2^-11 of the hi*hi one: its fp32 truncation matters 2^-11 as much, so it keeps
    // round 1's 1024 indices per drain and the epilogue converts half as many values on the other drains
    // NNLM_TC_HALVES=2 (experiment): 32 indices per accumulation for NP = 128, two chunks per k-block. Measured at config 5's full
    // size (scratch/c5_t1_full.py): T = 1 rel H 1.26e-5 with 64 indices AND with 32 (rel W 4.1e-9 -> 3.6e-9), at -10 % speed — what
    // is left there is the 22-24-bit storage of A itself (the oracle run on fp32-rounded A moves as much), so the default stays 1.
    static const int halves_env = [] { const char* e = getenv("NNLM_TC_HALVES"); return e ? atoi(e) : 0; }();
    p.halves = (MODE == 0 && NP >= 128 && p.drain == 1 && halves_env == 2) ? 2 : 1;
    p.d1_every = MODE == 0 ? std::max(1, 16 * p.halves / p.drain) : 1;
    p.center = center; p.fsum = fsum;
    kern<<<plan.grid, THREADS, smem, st>>>(mA_hi, mA_lo, mF_hi, mF_lo, p);
    NNLM_LAUNCHED();
}

}  // namespace

bool cross_tc_supported(int k) { return k >= 1 && k <= 128; }
int cross_tc_np(int k) { return k <= 32 ? 32 : (k <= 64 ? 64 : 128); }
int64_t cross_tc_ld(int64_t len) { return (len + 7) / 8 * 8; }     // TMA row pitch must be a multiple of 16 bytes

CrossPlan cross_tc_plan(int k, int64_t len, int64_t ncol, bool pairs)
{
    NNLM_REQUIRE(!pairs || k == 128, "the CTA-pair contraction runs on 128-row factor tiles");
    CrossPlan pl;
    pl.k = k; pl.len = len; pl.ncol = ncol; pl.pairs = pairs;
    pl.np = cross_tc_np(k);
    pl.ld_a = cross_tc_ld(len);
    pl.ld_f = cross_tc_ld(len);
    pl.tiles = ceil_div(ncol, pairs ? 2 * BM : BM);
    pl.kblocks = ceil_div(len, BK);
    pl.units = pl.tiles * pl.kblocks;
    // pairs: `grid` counts PAIRS of CTAs (the launch uses 2 * grid CTAs in clusters of two)
    pl.grid = (int)std::min<int64_t>(pairs ? sm_count() / 2 : sm_count(), pl.units);
    // most CTAs that touch one tile
    int64_t worst = 1;
    for (int64_t t = 0; t < pl.tiles; t++) {
        const int64_t first = first_cta_of_unit(t * pl.kblocks, pl.units, pl.grid);
        const int64_t last = first_cta_of_unit((t + 1) * pl.kblocks - 1, pl.units, pl.grid);
        worst = std::max(worst, last - first + 1);
    }
    pl.slots = (int)worst;
    return pl;
}

void launch_cross_tc(const CrossPlan& plan, const __half* a_hi, const __half* a_lo, const __half* f_hi, const __half* f_lo,
                     const double* unscale, const double* center, const double* fsum, double* Qp, cudaStream_t st, int drain)
{
    NNLM_REQUIRE(cross_tc_supported(plan.k), "tensor-core cross-product supports rank k <= 128");
    if (plan.np == 32)      launch_np<32, 5, 0>(plan, a_hi, a_lo, f_hi, f_lo, unscale, center, fsum, Qp, drain, st);
    else if (plan.np == 64) launch_np<64, 4, 0>(plan, a_hi, a_lo, f_hi, f_lo, unscale, center, fsum, Qp, drain, st);
    else                    launch_np<128, 3, 0>(plan, a_hi, a_lo, f_hi, f_lo, unscale, center, fsum, Qp, drain, st);
}

// Exact contraction of two integer-valued fp16 planes (MODE 1): Qp[slot][ncol][128] = unscale[a] * sum_i f[a,i] * a[i,j].
// |f| <= 2047 and a in {0, 1}: a TMEM partial sum over 64 k-blocks (4096 indices) stays below 2^24, hence exact in fp32.
void launch_cross_tc_exact(const CrossPlan& plan, const __half* a_plane, const __half* f_plane, const double* unscale, double* Qp,
                           cudaStream_t st)
{
    NNLM_REQUIRE(plan.np == 128, "the exact integer contraction runs on 128-row factor tiles");
    launch_np<128, 6, 1>(plan, a_plane, nullptr, f_plane, nullptr, unscale, nullptr, nullptr, Qp, 64, st);
}

// Two consecutive slices in one pass (MODE 2): Qp = unscale[a] * sum_i (f0[a,i] + f1[a,i] / 2048) * a[i,j], exact: both partial
// sums are integers below 2^24 and their fp64 combination spans 38 bits. unscale = the factor of slice f0.
void launch_cross_tc_exact2(const CrossPlan& plan, const __half* a_plane, const __half* f0, const __half* f1, const double* unscale,
                            double* Qp, cudaStream_t st)
{
    NNLM_REQUIRE(plan.np == 128, "the exact integer contraction runs on 128-row factor tiles");
    launch_np<128, 4, 2>(plan, a_plane, nullptr, f0, f1, unscale, nullptr, nullptr, Qp, 64, st);
}

// The same contraction by CTA pairs (k_mask_tc2); plan from cross_tc_plan(128, len, ncol, /*pairs=*/true).
void launch_mask_tc2(const CrossPlan& plan, const __half* a_plane, const __half* f0, const __half* f1, const double* unscale,
                     double* Qp, cudaStream_t st)
{
    NNLM_REQUIRE(plan.np == 128 && plan.pairs && plan.k == 128, "the CTA-pair contraction needs a pair plan of rank 128");
    constexpr size_t smem = (size_t)P2_STAGES * 2 * 128 * BK * 2 + 1024 /*alignment slack*/ + 256;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(k_mask_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const CUtensorMap mA = make_map(a_plane, plan.len, plan.ncol, plan.ld_a, 128);
    const CUtensorMap mF0 = make_map(f0, plan.len, 128, plan.ld_f, 128);
    const CUtensorMap mF1 = make_map(f1, plan.len, 128, plan.ld_f, 128);
    CrossParams p;
    p.ncol = plan.ncol; p.kblocks = plan.kblocks; p.units = plan.units; p.k = plan.k; p.slots = plan.slots; p.Qp = Qp; p.unscale = unscale;
    p.drain = 64; p.d1_every = 1; p.halves = 1; p.center = nullptr; p.fsum = nullptr;
    k_mask_tc2<<<2 * plan.grid, THREADS, smem, st>>>(mA, mF0, mF1, p);
    NNLM_LAUNCHED();
}

void launch_means(const double* A, int64_t len, int64_t ncol, double* colmean, double* rowmean, cudaStream_t st)
{
    if (colmean) {
        k_col_means<<<(int)std::min<int64_t>(ceil_div(ncol * 32, 256), 148 * 8), 256, 0, st>>>(A, len, ncol, colmean);
        NNLM_LAUNCHED();
    }
    if (rowmean) {
        k_row_means<<<(unsigned)ceil_div(len, 128), 128, 0, st>>>(A, len, ncol, rowmean);
        NNLM_LAUNCHED();
    }
}

void launch_absmax(const double* A, int64_t total, unsigned long long* maxbits, cudaStream_t st)
{
    NNLM_CUDA_CHECK(cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long), st));
    if (total <= 0) return;
    k_absmax<<<(int)std::min<int64_t>(ceil_div(total, 1024), 148 * 8), 1024, 0, st>>>(A, total, maxbits);
    NNLM_LAUNCHED();
}

void launch_scale_from_max(const unsigned long long* maxbits, double* sA, cudaStream_t st)
{
    k_scale_from_max<<<1, 1, 0, st>>>(maxbits, sA);
    NNLM_LAUNCHED();
}

void launch_split_matrix(const double* A, int64_t len, int64_t ncol, const double* sA, const double* colmean,
                         const double* rowmean, __half* a_hi, __half* a_lo, int64_t ld_a,
                         __half* t_hi, __half* t_lo, int64_t ld_t, cudaStream_t st)
{
    NNLM_REQUIRE(ceil_div(ncol, 32) <= 65535, "too many columns for the plane conversion grid");
    dim3 grid((unsigned)ceil_div(len, 32), (unsigned)ceil_div(ncol, 32));
    k_split_matrix<<<grid, 256, 0, st>>>(A, len, ncol, sA, colmean, rowmean, a_hi, a_lo, ld_a, t_hi, t_lo, ld_t);
    NNLM_LAUNCHED();
}

// Row sums and row maxima of the factor in ONE pass, finished by the last CTA to arrive (fixed-order reduction of the
// per-CTA partials, so the result does not depend on which CTA that is): fsum[a] = sum_i F[a,i] (the `sumW` of
// src/update_with_missing.cpp:27 and the mean-correction term of the cross-product), scales / unscale as k_make_scales.
// Replaces five launches of round 1 (memset, k_rowmax, k_make_scales, k_rowsum_partial, k_rowsum_finish).
__global__ void __launch_bounds__(256)
k_factor_prep(const double* __restrict__ F, int k, int64_t len, int64_t per_split, int np, const double* __restrict__ sA,
              double* __restrict__ part /* [splits][2][k] */, unsigned int* __restrict__ ticket, double* __restrict__ fsum,
              double* __restrict__ scales, double* __restrict__ unscale)
{
    __shared__ double sm[2][256];
    __shared__ bool s_last;
    const int groups = 256 / k;                                   // k <= 128
    const int a = threadIdx.x % k, g = threadIdx.x / k;
    const int64_t i_beg = (int64_t)blockIdx.x * per_split, i_end = min(len, i_beg + per_split);
    double s = 0.0, mx = 0.0;
    if (g < groups)
        for (int64_t i = i_beg + g; i < i_end; i += groups) {
            const double v = F[a + (int64_t)k * i];
            s += v;
            const double av = fabs(v);
            if (!is_missing(av)) mx = fmax(mx, av);
        }
    sm[0][threadIdx.x] = s; sm[1][threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x < k) {
        double t = 0.0, m2 = 0.0;
        for (int gg = 0; gg < groups; gg++) { t += sm[0][threadIdx.x + gg * k]; m2 = fmax(m2, sm[1][threadIdx.x + gg * k]); }
        part[((int64_t)blockIdx.x * 2 + 0) * k + threadIdx.x] = t;
        part[((int64_t)blockIdx.x * 2 + 1) * k + threadIdx.x] = m2;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int splits = gridDim.x;
    for (int r = warp; r < np; r += 8) {                          // one warp per row, lanes stride the partials in a fixed order
        double t = 0.0, m2 = 0.0;
        if (r < k)
            for (int sp = lane; sp < splits; sp += 32) {
                t += part[((int64_t)sp * 2 + 0) * k + r];
                m2 = fmax(m2, part[((int64_t)sp * 2 + 1) * k + r]);
            }
        t = warp_sum(t);
#pragma unroll
        for (int x = 16; x > 0; x >>= 1) m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, x));
        if (lane == 0) {
            const double sc = (r < k) ? pow2_scale(m2) : 1.0;
            if (r < k) fsum[r] = t;
            scales[r] = sc;
            unscale[r] = 1.0 / (sA[0] * sc);
        }
    }
    if (threadIdx.x == 0) *ticket = 0;                            // ready for the next half-iteration
}

void launch_rowmax(const double* F, int k, int64_t len, unsigned long long* rowmax, cudaStream_t st)
{
    NNLM_CUDA_CHECK(cudaMemsetAsync(rowmax, 0, sizeof(unsigned long long) * k, st));
    const int64_t total = (int64_t)k * len;
    k_rowmax<<<(int)std::min<int64_t>(ceil_div(total, 256 * 8), 148 * 4), 256, sizeof(unsigned long long) * k, st>>>(F, k, len, rowmax);
    NNLM_LAUNCHED();
}

void launch_factor_prep(const double* F, int k, int64_t len, int64_t ld, int np, const double* sA, double* part, int splits,
                        unsigned int* ticket, double* fsum, double* scales, double* unscale, __half* hi, __half* lo, cudaStream_t st)
{
    NNLM_REQUIRE(k >= 1 && k <= 128, "factor preparation supports k <= 128");
    const int64_t per_split = ceil_div(len, splits);
    k_factor_prep<<<splits, 256, 0, st>>>(F, k, len, per_split, np, sA, part, ticket, fsum, scales, unscale);
    NNLM_LAUNCHED();
    const size_t smem = sizeof(double) * 64 * (k + 1);
    if (smem > 48 * 1024) NNLM_CUDA_CHECK(cudaFuncSetAttribute(k_split_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_split_factor<<<(unsigned)ceil_div(ld, 64), 256, smem, st>>>(F, k, len, ld, np, scales, hi, lo);
    NNLM_LAUNCHED();
}

void launch_split_factor(const double* F, int k, int64_t len, int64_t ld, int np, const double* sA, unsigned long long* rowmax,
                         double* scales, double* unscale, __half* hi, __half* lo, cudaStream_t st)
{
    NNLM_CUDA_CHECK(cudaMemsetAsync(rowmax, 0, sizeof(unsigned long long) * np, st));
    const int64_t total = (int64_t)k * len;
    k_rowmax<<<(int)std::min<int64_t>(ceil_div(total, 256 * 8), 148 * 4), 256, sizeof(unsigned long long) * k, st>>>(F, k, len, rowmax);
    NNLM_LAUNCHED();
    k_make_scales<<<1, 128, 0, st>>>(rowmax, k, np, sA, scales, unscale);     // np <= 128
    NNLM_LAUNCHED();
    const size_t smem = sizeof(double) * 64 * (k + 1);
    if (smem > 48 * 1024) NNLM_CUDA_CHECK(cudaFuncSetAttribute(k_split_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_split_factor<<<(unsigned)ceil_div(ld, 64), 256, smem, st>>>(F, k, len, ld, np, scales, hi, lo);
    NNLM_LAUNCHED();
}

}  // namespace nnlm
