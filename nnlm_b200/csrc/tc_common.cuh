// tc_common.cuh — the sm_100a building blocks shared by the tensor-core kernels of this library (cross_tc.cu, error_tc.cu):
// mbarrier / TMA / tcgen05 PTX wrappers, the K-major SWIZZLE_128B operand descriptor, TMEM loads, the fp16 hi/lo split and
// the host-side tensor-map encoder. Written in-tree (no CUTLASS include).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <string>

#include "common.cuh"

namespace nnlm {
namespace tc {

constexpr float LO_SCALE = 2048.0f;            // 2^11: the lo plane holds (x - hi) * 2^11
constexpr double LO_UNSCALE = 1.0 / 2048.0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major SWIZZLE_128B operand descriptor: 8-row groups 1024 bytes apart (SBO), version 1 (sm_100), layout type 2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset
    d |= (uint64_t)1 << 46;                  // descriptor version
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}

template <int N> struct TmemLd;
template <> struct TmemLd<16> {
    __device__ static __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[16]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
    }
};
template <> struct TmemLd<32> {
    __device__ static __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
    }
};

// power-of-two scale that brings `maxabs` into [2^13, 2^14): fp16 keeps 11 significant bits there and the scaled
// remainder (< 2^3 before the 2^11 lift) stays far from both overflow and the subnormal range
__device__ __forceinline__ double pow2_scale(double maxabs)
{
    if (!(maxabs > 0.0)) return 1.0;
    int e;
    frexp(maxabs, &e);                 // maxabs = f * 2^e, f in [0.5, 1)
    return ldexp(1.0, 14 - e);
}

__device__ __forceinline__ void split2(double x, __half& hi, __half& lo)
{
    const float xf = (float)x;                         // |x| <= 2^14: conversion error 2^-24 relative, below the lo plane
    hi = __float2half_rn(xf);
    const double rem = x - (double)__half2float(hi);
    lo = __float2half_rn((float)(rem * (double)LO_SCALE));
}


// ---------------------------------------------------------------------------------------------------- host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeFn encode_fn()
{
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NNLM_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) throw Error(NNLM_E_CUDA, "cuTensorMapEncodeTiled is not available");
        fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

// 2-D fp16 tensor: inner extent `inner` (contiguous), `rows` rows of pitch ld elements; box = 64 x box_rows, SWIZZLE_128B
inline CUtensorMap make_map(const __half* base, int64_t inner, int64_t rows, int64_t ld, int box_rows)
{
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
    const cuuint32_t box[2] = {(cuuint32_t)64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw Error(NNLM_E_CUDA, "cuTensorMapEncodeTiled failed (rc " + std::to_string((int)rc) + ")");
    return m;
}

inline int sm_count()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}


}  // namespace tc
}  // namespace nnlm
