// norm.cuh — RMSNorm folded into the activation path of the Linear kernels (second half).
//
// Mila normalises, stores BF16, and the Linear reads that BF16 tensor back (Gemma.Block.ixx:209-210,347-349; kernel
// Normalizations/RmsNorm/Kernels/RmsNorm.Bf16.cu:19-73).  The fused path produces EXACTLY the BF16 values that kernel
// would have stored — same reduction order, same expression, same roundings — and hands them to the activation split
// in registers, so Linear(RMSNorm(x)) equals the two-kernel sequence bit for bit without the [M, K] round trip:
//   m2   = per lane i: fma chain over x[i], x[i + 32], ... ; then the shfl_down tree 16, 8, 4, 2, 1 (RmsNorm.Bf16.cu:48-58)
//   rstd = rsqrtf(m2 / float(K) + eps)                                                            (:60)
//   out  = bf16( fma(x * rstd, float(weight[k]) + weight_offset, bias ? float(bias[k]) : 0.0f) )    (:68-71; nvcc contracts
//          `x * rstd * w + b` into one FMUL and one FFMA, also when b is the literal 0.0f)
#pragma once
#include "common.cuh"

namespace milab200 {

struct NormArgs {
    const __nv_bfloat16* weight = nullptr;     // [K] or null (w = 1)
    const __nv_bfloat16* bias = nullptr;       // [K] or null
    float eps = 0.0f, weight_offset = 0.0f;
    int on = 0;
    int fast = 0;                              // opt-in (milab200_set_option "rmsnorm_fast_reduction"): sum of squares in tree order
                                               //   with 128-bit loads instead of the reference's lane-strided FMA chains — rstd within a
                                               //   few FP32 ulps of the reference's, ~0.5 us instead of ~3 us on the dependency chain of a
                                               //   decode Linear; the fused result is then no longer bit-identical to the two-kernel sequence
};

// One warp, one token row x[0..K): the reference's rstd, bit for bit.  Every lane returns it.
__device__ __forceinline__ float rms_rstd_warp(const __nv_bfloat16* __restrict__ x, int K, float eps, int lane)
{
    float m2 = 0.0f;
    // The lane's FMA chain is sequential by definition, its loads are not: this runs on the dependency chain of a decode
    // Linear (previous kernel's last store -> rstd -> split -> first MMA).  32 loads in flight per L2 round trip, the last
    // (partial) batch guarded; same order of operations, same bits.  (A deeper batch costs the host kernels registers,
    // and as a non-inlined function every load pays two R2UR: both measured slower, profiles/r2k10_r2k13_*.)
    int nsteps = (K - lane + 31) >> 5;                           // elements of this lane's chain: x[lane], x[lane + 32], ...
    const unsigned short* pb = reinterpret_cast<const unsigned short*>(x) + lane;
    for (; nsteps >= 32; nsteps -= 32, pb += 32 * 32) {
        unsigned short raw[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) raw[u] = pb[32 * u];
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const float v = __uint_as_float((unsigned)raw[u] << 16);
            m2 = fmaf(v, v, m2);
        }
    }
    if (nsteps > 0) {
        unsigned short raw[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) raw[u] = (u < nsteps) ? pb[32 * u] : (unsigned short)0;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const float v = __uint_as_float((unsigned)raw[u] << 16);                           // past the end: fma(0, 0, m2) == m2
            m2 = fmaf(v, v, m2);
        }
    }
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) m2 += __shfl_down_sync(0xffffffffu, m2, offset);
    m2 = __shfl_sync(0xffffffffu, m2, 0);
    return rsqrtf(__fdiv_rn(m2, (float)K) + eps);
}

// Tree-order variant (NormArgs::fast): K % 8 == 0.  Every lane sums the squares of 8 consecutive elements per 128-bit load
// (8 loads in flight), then a butterfly over the warp.  Not the reference's order of operations: the result differs from
// rms_rstd_warp in the last FP32 bits.
__device__ __forceinline__ float rms_rstd_warp_fast(const __nv_bfloat16* __restrict__ x, int K, float eps, int lane)
{
    const uint4* xr = reinterpret_cast<const uint4*>(x);
    const int n8 = K >> 3;
    float m2 = 0.0f;
    for (int i = lane; i < n8; i += 32 * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + 32 * u < n8) ? xr[i + 32 * u] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t w[4] = { v[u].x, v[u].y, v[u].z, v[u].w };
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16lo(w[j]), b = bf16hi(w[j]);
                m2 = fmaf(a, a, m2);
                m2 = fmaf(b, b, m2);
            }
        }
    }
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, offset);
    return rsqrtf(__fdiv_rn(m2, (float)K) + eps);
}

// the reciprocal RMS the fused routes use: the reference's order unless the caller opted into the tree-order reduction
__device__ __forceinline__ float rms_rstd_select(const NormArgs& n, const __nv_bfloat16* __restrict__ x, int K, int lane)
{
    if (n.fast && (K & 7) == 0) return rms_rstd_warp_fast(x, K, n.eps, lane);
    return rms_rstd_warp(x, K, n.eps, lane);
}

__device__ __forceinline__ float rms_apply1(float xv, float rstd, float w, float b)
{
    return fmaf(__fmul_rn(xv, rstd), w, b);
}

// 8 consecutive BF16 activations (one uint4) of a token with reciprocal RMS `rstd`, their 8 norm weights / biases
// (uint4 of BF16, ignored when the pointer they came from was null) -> the 8 BF16 values RMSNorm would have stored.
__device__ __forceinline__ uint4 rms_apply8(const uint4& x8, float rstd, const uint4& w8, const uint4& b8, bool has_w, bool has_b,
                                            float weight_offset)
{
    const uint32_t xw[4] = { x8.x, x8.y, x8.z, x8.w }, ww[4] = { w8.x, w8.y, w8.z, w8.w }, bw[4] = { b8.x, b8.y, b8.z, b8.w };
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float w0 = has_w ? __fadd_rn(bf16lo(ww[j]), weight_offset) : 1.0f, w1 = has_w ? __fadd_rn(bf16hi(ww[j]), weight_offset) : 1.0f;
        const float b0 = has_b ? bf16lo(bw[j]) : 0.0f, b1 = has_b ? bf16hi(bw[j]) : 0.0f;
        const __nv_bfloat16 r0 = __float2bfloat16(rms_apply1(bf16lo(xw[j]), rstd, w0, b0));
        const __nv_bfloat16 r1 = __float2bfloat16(rms_apply1(bf16hi(xw[j]), rstd, w1, b1));
        o[j] = (uint32_t)__bfloat16_as_ushort(r0) | ((uint32_t)__bfloat16_as_ushort(r1) << 16);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace milab200
