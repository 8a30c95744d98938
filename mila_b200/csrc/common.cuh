// common.cuh — device helpers shared by the sm_100a kernels of libmila_b200_linear.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "../../include/mila_b200_linear.h"

namespace milab200 {

// ---- launch accounting (host) -----------------------------------------------------------
void note_launch(const char* kernel_name, uint64_t n = 1);

#define MILAB200_RETURN_IF_CUDA(expr)                         \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) return (int)_e; } while (0)

// ---- memory ----------------------------------------------------------------------------

// Streaming 128-bit load: weights are read exactly once, keep them out of L1.
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
// Cached 128-bit load (activations: re-read by every CTA, L1/L2 resident).
__device__ __forceinline__ uint4 ldg_cached_v4(const void* p)
{
    return __ldg(reinterpret_cast<const uint4*>(p));
}

// ---- scalar conversions ------------------------------------------------------------------

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h16) { return __uint_as_float(h16 << 16); }
__device__ __forceinline__ float bf16lo(uint32_t pair) { return __uint_as_float(pair << 16); }
__device__ __forceinline__ float bf16hi(uint32_t pair) { return __uint_as_float(pair & 0xFFFF0000u); }

__device__ __forceinline__ uint32_t pack_f16x2_rn(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// 4 packed E4M3 bytes -> two f16x2 (exact).  lo = bytes 0,1 ; hi = bytes 2,3.
__device__ __forceinline__ void cvt_e4m3x4_to_f16x2x2(uint32_t w, uint32_t& lo, uint32_t& hi)
{
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2;\n"
        "  cvt.rn.f16x2.e4m3x2 %0, l; cvt.rn.f16x2.e4m3x2 %1, h; }"
        : "=r"(lo), "=r"(hi) : "r"(w));
}

// 8 packed E2M1 nibbles (one 32-bit word, low nibble = even element) -> four f16x2 (exact);
// o[b] = elements (2b, 2b+1) of the word.  One F2FP.F16.E2M1.UNPACK_B per pair on sm_100a.
__device__ __forceinline__ void cvt_e2m1x8_to_f16x2x4(uint32_t w, uint32_t (&o)[4])
{
    asm("{ .reg .b8 a, b, c, d; mov.b32 {a, b, c, d}, %4;\n"
        "  cvt.rn.f16x2.e2m1x2 %0, a; cvt.rn.f16x2.e2m1x2 %1, b;\n"
        "  cvt.rn.f16x2.e2m1x2 %2, c; cvt.rn.f16x2.e2m1x2 %3, d; }"
        : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(w));
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_m16n8k16_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                                 uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 "
                 "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// FP4 E2M1 magnitude LUT decode (scalar paths only).
__device__ __forceinline__ float e2m1_to_f32(uint32_t nib)
{
    // {0, .5, 1, 1.5, 2, 3, 4, 6}: exponent/mantissa arithmetic instead of a memory LUT
    const uint32_t c = nib & 7u;
    // c<2 -> c*0.5 ; else 2^((c>>1)-1) * (1 + (c&1)/2)
    float mag = (c < 2u) ? 0.5f * (float)c
                         : __uint_as_float(((126u + (c >> 1)) << 23) | ((c & 1u) << 22));
    return (nib & 8u) ? -mag : mag;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

}  // namespace milab200
