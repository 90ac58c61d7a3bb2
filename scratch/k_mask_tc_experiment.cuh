// Experiment (round 2, not compiled into the library): 256-column tiles for the exact mask contraction, single-buffered TMEM.
// Slower than k_cross_tc<128,4,2> (331 vs 249 us under ncu): see profiles/r2_d_config4.md, 'What bounds the contraction'.
// Fragment of cross_tc.cu (uses its helpers).
// ---------------------------------------------------------------------------------------------------- wide exact contraction
// k_mask_tc: the exact integer contraction of the NA path (MODE 2's arithmetic) on a 256 x 128 x 2-slice tile per CTA.
// Why: MODE 2 moves 48 KB through L2 per k-block for 4.2 Mflop (87 flop/B); at the chip's L2 throughput (~6300 B/clk, 12.4 TB/s)
// that caps the tensor pipe at ~1080 TFLOP/s, and the kernel sat exactly there (992 TFLOP/s measured, profiles/r2_d_config4.md).
// With TWO 128-row tiles of the mask against the same two slice tiles a k-block moves 64 KB for 8.4 Mflop (131 flop/B).
// The four fp32 accumulators (2 mask tiles x 2 slices x 128 columns) fill TMEM, so there is no double buffering: the MMA
// issuer waits while the epilogue drains, once per 4096 contraction indices (exactness bound of the fp32 partial sums) — a few
// microseconds per CTA and launch. The epilogue keeps no accumulator in registers: every drain adds into the CTA's own slot
// of the partial buffer in global memory (the first drain of a tile segment stores).
constexpr int XM = 256;                                  // mask columns per tile
constexpr int X_EPI_WARPS = 8;
constexpr int X_THREADS = 32 * (2 + X_EPI_WARPS);
constexpr int X_STAGES = 3;

__global__ void __launch_bounds__(X_THREADS, 1)
k_mask_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapF0,
          const __grid_constant__ CUtensorMap mapF1, const CrossParams p)
{
    constexpr int NP = 128;
    constexpr int T_BYTES = 128 * BK * 2;                  // 16 KB: one 128-row operand tile per stage
    constexpr int STAGE_BYTES = 4 * T_BYTES;               // mask rows 0..127, mask rows 128..255, slice 0, slice 1
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + X_STAGES * STAGE_BYTES);
    uint64_t* empty = full + X_STAGES;
    uint64_t* tfull = empty + X_STAGES;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    double* s_us = reinterpret_cast<double*>(smem + X_STAGES * STAGE_BYTES + 128);       // [128] unscale factors

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t P = gridDim.x, U = p.units, KBn = p.kblocks;
    const int64_t u0 = ((int64_t)blockIdx.x * U) / P, u1 = ((int64_t)(blockIdx.x + 1) * U) / P;

    if (threadIdx.x < 128) s_us[threadIdx.x] = p.unscale[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < X_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1); mbar_init(tempty, X_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t u = u0; u < u1; u++) {
                const int64_t tile = u / KBn, kb = u % KBn;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = tiles + stage * STAGE_BYTES;
                mbar_expect_tx(&full[stage], STAGE_BYTES);
                const int c0 = (int)(kb * BK), c1 = (int)(tile * XM);
                tma_load_2d(sa, &mapA, &full[stage], c0, c1);
                tma_load_2d(sa + T_BYTES, &mapA, &full[stage], c0, c1 + 128);      // (rows past the end read as zero)
                tma_load_2d(sa + 2 * T_BYTES, &mapF0, &full[stage], c0, 0);
                tma_load_2d(sa + 3 * T_BYTES, &mapF1, &full[stage], c0, 0);
                if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t chunk = 0;
            int64_t u = u0;
            while (u < u1) {
                const int64_t tile = u / KBn;
                const int64_t seg_end = min(u1, (tile + 1) * KBn);
                while (u < seg_end) {
                    const int64_t chunk_end = min(seg_end, u + (int64_t)p.drain);
                    mbar_wait(tempty, (chunk & 1) ^ 1);            // the epilogue has drained the accumulators
                    tc_fence_after();
                    bool first = true;
                    for (; u < chunk_end; u++) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
                        const uint64_t a0 = make_desc(sa), a1 = make_desc(sa + T_BYTES);
                        const uint64_t f0 = make_desc(sa + 2 * T_BYTES), f1 = make_desc(sa + 3 * T_BYTES);
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ks++) {
                            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                            const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                            umma_f16(tmem_base + 0 * NP, a0 + adv, f0 + adv, IDESC, acc);
                            umma_f16(tmem_base + 1 * NP, a0 + adv, f1 + adv, IDESC, acc);
                            umma_f16(tmem_base + 2 * NP, a1 + adv, f0 + adv, IDESC, acc);
                            umma_f16(tmem_base + 3 * NP, a1 + adv, f1 + adv, IDESC, acc);
                        }
                        first = false;
                        tc_commit(&empty[stage]);
                        if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull);
                    chunk++;
                }
            }
        }
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int cg = ew >> 2;                       // which 64 of the 128 accumulator columns
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t chunk = 0;
        int64_t u = u0;
        while (u < u1) {
            const int64_t tile = u / KBn;
            const int64_t seg_end = min(u1, (tile + 1) * KBn);
            const int64_t slot = (int64_t)blockIdx.x - first_cta_of_unit(tile * KBn, U, P);
            bool first_chunk = true;
            while (u < seg_end) {
                const int64_t chunk_end = min(seg_end, u + (int64_t)p.drain);
                mbar_wait(tfull, chunk & 1);
                tc_fence_after();
#pragma unroll
                for (int mh = 0; mh < 2; mh++) {
                    const int64_t j = tile * XM + mh * 128 + quarter * 32 + lane;
#pragma unroll
                    for (int ch = 0; ch < 2; ch++) {
                        const int col = cg * 64 + ch * 32;
                        uint32_t r0[32], r1[32];
                        TmemLd<32>::ld(tmem_base + lane_addr + (2 * mh) * NP + col, r0);
                        TmemLd<32>::ld(tmem_base + lane_addr + (2 * mh + 1) * NP + col, r1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (j < p.ncol) {
                            double* out = p.Qp + (slot * p.ncol + j) * p.k + col;
                            if (first_chunk) {
#pragma unroll
                                for (int c = 0; c < 32; c++)
                                    out[c] = fma((double)__uint_as_float(r1[c]), LO_UNSCALE, (double)__uint_as_float(r0[c])) * s_us[col + c];
                            } else {
#pragma unroll
                                for (int c = 0; c < 32; c++)     // (a reduction without return value: the thread does not wait for the old value)
                                    atomicAdd(out + c, fma((double)__uint_as_float(r1[c]), LO_UNSCALE, (double)__uint_as_float(r0[c])) * s_us[col + c]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty);
                first_chunk = false;
                u = chunk_end;
                chunk++;
            }
            if (seg_end == (tile + 1) * KBn) {            // the CTA that finishes a tile zero-fills the slots nobody writes
#pragma unroll
                for (int mh = 0; mh < 2; mh++) {
                    const int64_t j = tile * XM + mh * 128 + quarter * 32 + lane;
                    if (j < p.ncol)
                        for (int64_t sp = slot + 1; sp < p.slots; sp++) {
                            double* z = p.Qp + (sp * p.ncol + j) * p.k + cg * 64;
#pragma unroll
                            for (int c = 0; c < 64; c++) z[c] = 0.0;
                        }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

