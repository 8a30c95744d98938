// act_split.cuh — exact split of BF16 activations into two E4M3 planes (the B operand of the
// tcgen05 kind::f8f6f4 kernels, decode_tc.cu and prefill_tc.cu) and the tensor-map encoder both use.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace milab200 {

// 8 consecutive BF16 activations (one uint4) -> 8 hi + 8 lo E4M3 bytes for the block scale inv = 2^-e.
// v = x * inv is exact in FP16 (8-bit significand, |v| <= 256); hi = rn_e4m3(v); lo = rn_e4m3(16 * (v - hi)),
// the subtraction and the scaling being exact in FP16.  Ten instructions per pair of activations.
// LOGAIN = 16: the lo plane is rn_e4m3(16 (v - hi)) and the consumer computes D_hi + D_lo / 16 (exact for |v| >= 2^-6).
// LOGAIN = 1 ("summed planes", batched FP4 path): lo = rn_e4m3(v - hi), so hi + lo = v for every |v| >= 2^-2 and the
// tensor core itself adds the planes (both accumulate into the same TMEM columns); below that the absolute error is
// < 2^-17 of the token maximum.
template <int LOGAIN>
__device__ __forceinline__ void split_e4m3x8_t(const uint4& v, float inv, uint2& hi, uint2& lo)
{
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
    uint16_t h[4], l[4];
    const __half2 k16 = __floats2half2_rn((float)LOGAIN, (float)LOGAIN);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x0 = bf16lo(w[j]) * inv, x1 = bf16hi(w[j]) * inv;
        uint32_t v16;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(v16) : "f"(x1), "f"(x0));
        asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(h[j]) : "r"(v16));
        // hi back to f16x2 without a second FP8 conversion (the F2FP pipe is the converter's bottleneck at M > 8):
        // an E4M3 byte shifted to bits [14:7] of a half IS that value times 2^-8 (same subnormal semantics), so
        // place the two bytes, fix the sign bits and multiply by 256 — exact.
        const uint32_t w8 = __byte_perm((uint32_t)h[j], 0u, 0x1404);                 // b0 << 8 | b1 << 24
        const uint32_t hs = ((w8 >> 1) & 0x3F803F80u) | (w8 & 0x80008000u);
        const __half2 k256 = __floats2half2_rn(256.0f, 256.0f);
        const __half2 hb = __hmul2(*reinterpret_cast<const __half2*>(&hs), k256);
        const uint32_t hb16 = *reinterpret_cast<const uint32_t*>(&hb);
        const __half2 d = __hmul2(__hsub2(*reinterpret_cast<const __half2*>(&v16), *reinterpret_cast<const __half2*>(&hb16)), k16);
        asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(l[j]) : "r"(*reinterpret_cast<const uint32_t*>(&d)));
    }
    hi = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    lo = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

__device__ __forceinline__ void split_e4m3x8(const uint4& v, float inv, uint2& hi, uint2& lo) { split_e4m3x8_t<16>(v, inv, hi, lo); }

// Inf/NaN activations poison their output row, as they would in FP32: force the E4M3 NaN code.
__device__ __forceinline__ void poison_nonfinite(const uint4& v, uint2& hi)
{
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
    uint32_t h[2] = { hi.x, hi.y };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if ((w[j] & 0x00007F80u) == 0x00007F80u) h[j >> 1] |= 0x7Fu << ((j & 1) * 16);
        if ((w[j] & 0x7F800000u) == 0x7F800000u) h[j >> 1] |= 0x7Fu << ((j & 1) * 16 + 8);
    }
    hi = make_uint2(h[0], h[1]);
}

// ---- host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

}  // namespace milab200
