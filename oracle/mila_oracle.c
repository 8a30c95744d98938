/*
 * mila_oracle.c — CPU restatement of Mila's quantized Linear<TWeightQuant> path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path is
 * the CUDA library (libmila_b200_linear.so); nothing in mila_b200/ links or calls this.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/Mila/Src/Dnn/, LIN = Compute/Devices/Cuda/Operations/Linear).
 *
 * Parity pins (see DESIGN.md "Oracle"):
 *   - E4M3 conversion: checked against the CUDA toolkit's own host implementation
 *     __nv_cvt_float_to_fp8 (cuda_fp8.hpp, the function Mila's kernel calls) by
 *     oracle/pin_cuda_fp8.cpp — exhaustive over all 2^32 floats once, sampled in tests.
 *   - Packed bytes / scales: checked on the GPU box against the reference's own kernels
 *     compiled unmodified from /root/reference into oracle/_ref/libmila_ref_linear.so.
 *   - The reference's tests hold no byte-level golden vectors; the
 *     formula pins they do hold (scale = absmax/448, decodeFp8E4M3, weightValue
 *     fixtures) are reproduced in tests/test_oracle.py.
 *
 * Plain C99; build: gcc -O2 -fPIC -shared (NO -ffast-math: IEEE division order matters).
 */

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>

/* ------------------------------------------------------------------------- */
/* scalar format helpers                                                     */
/* ------------------------------------------------------------------------- */

static inline float u32_as_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f32_as_u32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* BF16 -> FP32 is exact (__bfloat162float). */
float oracle_bf16_to_f32(uint16_t h) { return u32_as_f32((uint32_t)h << 16); }

/* FP32 -> BF16 round-to-nearest-even, NaN -> 0x7FFF (matches __float2bfloat16 /
 * cvt.rn.bf16.f32; same bit trick as the reference test fixture at
 * Tests/Dnn/Components/Linear/Linear.Cuda.cpp:674-676 for finite values). */
uint16_t oracle_f32_to_bf16(float f)
{
    uint32_t u = f32_as_u32(f);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FFFu;
    uint32_t rounding = 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)((u + rounding) >> 16);
}

/* FP32 -> FP8 E4M3, round-to-nearest-even, saturate-to-finite (±448 = 0x7E), NaN -> 0x7F.
 * Restates the semantics of cvt.rn.satfinite.e4m3x2.f32, which is what
 * `__nv_fp8_e4m3( float )` compiles to in LIN/Kernels/Quantization/
 * CudaFp8WeightQuantization.cu:120.  Written in FP32 bit arithmetic (the toolkit's host
 * version goes through double); pinned against it by pin_cuda_fp8.cpp. */
uint8_t oracle_f32_to_e4m3(float f)
{
    uint32_t u = f32_as_u32(f);
    uint32_t a = u & 0x7FFFFFFFu;
    uint8_t sign = (uint8_t)((u >> 24) & 0x80u);

    if (a > 0x7F800000u) return 0x7Fu;               /* NaN: canonical, sign dropped */
    /* 464 = 448 + half-ulp(32) is the first value that would round past 448. Everything
     * >= 464 (and inf) saturates; (448,464) rounds down to 448 under RNE?  448 = 1.75*2^8,
     * next grid point would be 480 = 1.875*2^8 (mantissa 111 = NaN slot), midpoint 464 ties
     * to even -> 480 slot -> saturates.  So: a >= 464 -> 0x7E; below handled by RNE + clamp. */
    if (a >= 0x43E80000u) return (uint8_t)(sign | 0x7Eu);   /* 464.0f = 0x43E80000 */

    int32_t e = (int32_t)(a >> 23) - 127;            /* unbiased exponent */
    uint32_t m = a & 0x007FFFFFu;

    if (e >= -6) {
        /* normal in E4M3: keep 3 mantissa bits, RNE on the 20 dropped bits */
        uint32_t keep = m >> 20;
        uint32_t rest = m & 0xFFFFFu;
        uint32_t r = ((uint32_t)(e + 7) << 3) | keep;
        if (rest > 0x80000u || (rest == 0x80000u && (keep & 1u))) r += 1u;
        if (r > 0x7Eu) r = 0x7Eu;                    /* cannot happen below 464, kept as guard */
        return (uint8_t)(sign | r);
    }
    /* subnormal in E4M3: value = q * 2^-9, q in 0..7 (q==8 rolls into min normal 0x08) */
    if (e < -10) {
        /* |x| < 2^-10 = half of min subnormal: rounds to zero (tie at exactly 2^-10 -> even -> 0) */
        return sign;
    }
    {
        /* full significand (1.m) as 24-bit integer, value = sig * 2^(e-23).
         * q = value / 2^-9 = sig * 2^(e-23+9) = sig >> (14 - e)  with e in [-10,-7]. */
        uint32_t sig = m | 0x00800000u;
        uint32_t shift = (uint32_t)(14 - e);         /* 21..24 */
        uint32_t q = sig >> shift;
        uint32_t rest = sig & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1u);
        if (rest > half || (rest == half && (q & 1u))) q += 1u;
        return (uint8_t)(sign | q);
    }
}

/* FP8 E4M3 -> FP32 (exact).  Same decode the reference's own test uses to check
 * reconstruction (Tests/Dnn/Components/Linear/Linear.Cuda.cpp:929-949). */
float oracle_e4m3_to_f32(uint8_t b)
{
    uint32_t e = (b >> 3) & 0xFu, m = b & 7u;
    float v;
    if (e == 0xFu && m == 7u) return u32_as_f32(0x7FFFFFFFu);
    if (e == 0) v = ldexpf((float)m, -9);
    else        v = ldexpf(1.0f + (float)m / 8.0f, (int)e - 7);
    return (b & 0x80u) ? -v : v;
}

/* FP4 E2M1 nibble encoder: strict-less-than threshold ladder, sign test `x < 0`.
 * Restates fp4_e2m1_quantize, LIN/Kernels/Quantization/CudaFp4WeightQuantization.cu:54-70.
 * Round-half-AWAY in magnitude (not RNE); -0.0f -> 0; NaN -> 7. */
uint8_t oracle_f32_to_e2m1(float x)
{
    uint8_t sign = (x < 0.0f) ? 8u : 0u;
    float a = fabsf(x);
    uint8_t mag;
    if      (a < 0.25f) mag = 0;
    else if (a < 0.75f) mag = 1;
    else if (a < 1.25f) mag = 2;
    else if (a < 1.75f) mag = 3;
    else if (a < 2.5f)  mag = 4;
    else if (a < 3.5f)  mag = 5;
    else if (a < 5.0f)  mag = 6;
    else                mag = 7;
    return (uint8_t)(sign | mag);
}

/* FP4 E2M1 nibble decode.  Restates fp4_e2m1_decode,
 * LIN/Kernels/MatVec/CudaMatVecBias.Bf16.cu:20-25 and Quantization/Weight/Policies.ixx:89-95. */
float oracle_e2m1_to_f32(uint8_t nib)
{
    static const float lut[8] = { 0.0f, 0.5f, 1.0f, 1.5f, 2.0f, 3.0f, 4.0f, 6.0f };
    float mag = lut[nib & 7u];
    return (nib & 8u) ? -mag : mag;
}

/* fmaxf as CUDA defines it (NaN operand is ignored) — glibc fmaxf has the same rule. */
static inline float maxf_nan_ignoring(float a, float b) { return fmaxf(a, b); }

/* ------------------------------------------------------------------------- */
/* load-time quantizers                                                      */
/* ------------------------------------------------------------------------- */

/* PerChannelFp8<FP8_E4M3>.  Restates quantize_fp8_per_channel_kernel,
 * LIN/Kernels/Quantization/CudaFp8WeightQuantization.cu:57-121:
 *   absmax = max_k |f32(W[n,k])| (fmaxf from 0 — NaNs ignored)
 *   scale  = absmax > 0 ? absmax / 448 : 1 ; inv = 1 / scale   (two IEEE fp32 divisions)
 *   W8[n,k] = e4m3_satfinite_rn( f32(W[n,k]) * inv )
 * src: BF16 bits [N,K] row-major.  dst: [N,K] bytes.  scales: [N] f32. */
void oracle_quantize_fp8_per_channel(const uint16_t* src, uint8_t* dst, float* scales,
                                     int64_t N, int64_t K)
{
    for (int64_t n = 0; n < N; ++n) {
        const uint16_t* row = src + n * K;
        float absmax = 0.0f;
        for (int64_t k = 0; k < K; ++k)
            absmax = maxf_nan_ignoring(absmax, fabsf(oracle_bf16_to_f32(row[k])));
        volatile float scale = (absmax > 0.0f) ? (absmax / 448.0f) : 1.0f;
        volatile float inv = 1.0f / scale;
        scales[n] = scale;
        for (int64_t k = 0; k < K; ++k) {
            volatile float v = oracle_bf16_to_f32(row[k]) * inv;
            dst[n * K + k] = oracle_f32_to_e4m3(v);
        }
    }
}

/* PerGroupFp4<g>, g in {64,128}.  Restates quantize_fp4_per_group_kernel,
 * LIN/Kernels/Quantization/CudaFp4WeightQuantization.cu:83-144:
 *   per (row, group): absmax (fmaxf tree over |v|), scale = absmax > 0 ? absmax/6 : 1,
 *   inv = 1/scale, nibble = e2m1(v * inv), byte b = nib(W[n,2b]) | nib(W[n,2b+1]) << 4.
 * Returns 0 on success, -1 for an unsupported group size / K % g != 0 (the reference
 * throws std::runtime_error for the former, :220). */
int oracle_quantize_fp4_per_group(const uint16_t* src, uint8_t* dst_packed, float* scales,
                                  int64_t N, int64_t K, int group_size)
{
    if (group_size != 64 && group_size != 128) return -1;
    if (K % group_size != 0) return -1;
    const int64_t G = K / group_size;
    for (int64_t n = 0; n < N; ++n) {
        const uint16_t* row = src + n * K;
        uint8_t* prow = dst_packed + n * (K / 2);
        for (int64_t g = 0; g < G; ++g) {
            float absmax = 0.0f;
            for (int i = 0; i < group_size; ++i)
                absmax = maxf_nan_ignoring(absmax, fabsf(oracle_bf16_to_f32(row[g * group_size + i])));
            /* note: the reference seeds the tree with |v| itself (not 0); with all-NaN
             * groups smem[0] is NaN there -> `absmax > 0` false -> scale 1.  fmaxf from 0
             * gives absmax 0 -> scale 1: same result. */
            volatile float scale = (absmax > 0.0f) ? (absmax / 6.0f) : 1.0f;
            volatile float inv = 1.0f / scale;
            scales[n * G + g] = scale;
            for (int i = 0; i < group_size; i += 2) {
                volatile float v0 = oracle_bf16_to_f32(row[g * group_size + i]) * inv;
                volatile float v1 = oracle_bf16_to_f32(row[g * group_size + i + 1]) * inv;
                uint8_t n0 = oracle_f32_to_e2m1(v0);
                uint8_t n1 = oracle_f32_to_e2m1(v1);
                prow[(g * group_size + i) / 2] = (uint8_t)(n0 | (n1 << 4));
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* dequantisation (FP32-exact weights)                                       */
/* ------------------------------------------------------------------------- */

/* w = f32(W8) * scale[n]   (LIN/Kernels/MatVec/CudaMatVecBias.Bf16.cu:234-249 applies the
 * scale after the K-sum; the dequantise-then-GEMM oracle applies it per element). */
void oracle_dequant_fp8(const uint8_t* w8, const float* scales, float* out, int64_t N, int64_t K)
{
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k)
            out[n * K + k] = oracle_e4m3_to_f32(w8[n * K + k]) * scales[n];
}

/* w = lut[nib] * scale[n, k/g]   (Bf16.cu:20-25, :318-326; Policies.ixx:89-95). */
int oracle_dequant_fp4(const uint8_t* packed, const float* scales, float* out,
                       int64_t N, int64_t K, int group_size)
{
    if (group_size != 64 && group_size != 128) return -1;
    if (K % group_size != 0) return -1;
    const int64_t G = K / group_size;
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) {
            uint8_t byte = packed[n * (K / 2) + k / 2];
            uint8_t nib = (k & 1) ? (uint8_t)(byte >> 4) : (uint8_t)(byte & 0xFu);
            out[n * K + k] = oracle_e2m1_to_f32(nib) * scales[n * G + k / group_size];
        }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* forward oracles                                                           */
/* ------------------------------------------------------------------------- */

/* The north star's result oracle: "dequantise-then-FP32-GEMM".
 *   Y[m,n] = bf16( sum_k f32(X[m,k]) * Wf[n,k] + f32(bias[n]) )
 * Wf = FP32-exact dequantised weights (functions above).  The sum is carried in double
 * so the oracle itself adds no ordering noise (the reference's own host check uses a
 * double accumulator too: Tests/.../Linear.Cuda.cpp:84-105); it is also the semantic of
 * the reference's per-row matvec loop (LIN/CudaLinearOp.ixx:833-875).
 * x: BF16 bits [M,K]; wf: f32 [N,K]; bias: BF16 bits [N] or NULL; y: BF16 bits [M,N];
 * y_f32 (optional, may be NULL): the unrounded result, for tolerance arithmetic. */
void oracle_linear_forward_bf16(const uint16_t* x, const float* wf, const uint16_t* bias,
                                uint16_t* y, float* y_f32, int64_t M, int64_t K, int64_t N)
{
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            double acc = 0.0;
            const uint16_t* xr = x + m * K;
            const float* wr = wf + n * K;
            for (int64_t k = 0; k < K; ++k)
                acc += (double)oracle_bf16_to_f32(xr[k]) * (double)wr[k];
            if (bias) acc += (double)oracle_bf16_to_f32(bias[n]);
            float r = (float)acc;
            if (y_f32) y_f32[m * N + n] = r;
            y[m * N + n] = oracle_f32_to_bf16(r);
        }
}

/* CpuLinearOp::forwardNaive — Compute/Devices/Cpu/Operations/CpuLinearOp.ixx:384-411.
 * Y = X * W^T + B, FP32 in/out, `long double` accumulator (x87 80-bit on Linux/GCC),
 * single thread at batch <= 100 (OpenMP is off by default and gated on batch > 100, :229).
 * This is the reference's CPU path for BASELINE.json config #1 and the cpu_baseline leg. */
void oracle_cpu_linear_forward_naive(const float* X, float* Y, const float* W, const float* B,
                                     int64_t batch, int64_t in_features, int64_t out_features)
{
    for (int64_t idx = 0; idx < batch; ++idx) {
        const int64_t in_base = idx * in_features;
        const int64_t out_base = idx * out_features;
        for (int64_t o = 0; o < out_features; ++o) {
            long double acc = 0.0L;
            for (int64_t i = 0; i < in_features; ++i)
                acc += (long double)X[in_base + i] * (long double)W[o * in_features + i];
            if (B) acc += (long double)B[o];
            Y[out_base + o] = (float)acc;
        }
    }
}

/* CpuLinearOp::forwardUnrolled — CpuLinearOp.ixx:418-456 (taken when batch % 8 == 0):
 * float accumulators seeded with the bias, 8 batch rows at a time. */
void oracle_cpu_linear_forward_unrolled(const float* X, float* Y, const float* W, const float* B,
                                        int64_t batch, int64_t in_features, int64_t out_features)
{
    enum { U = 8 };
    for (int64_t out_idx = 0; out_idx + U <= batch; out_idx += U)
        for (int64_t o = 0; o < out_features; ++o) {
            float r[U];
            for (int j = 0; j < U; ++j) r[j] = B ? B[o] : 0.0f;
            for (int64_t i = 0; i < in_features; ++i) {
                float w = W[o * in_features + i];
                for (int j = 0; j < U; ++j) {
                    volatile float p = X[(out_idx + j) * in_features + i] * w; /* no FMA contraction */
                    r[j] += p;
                }
            }
            for (int j = 0; j < U; ++j) Y[(out_idx + j) * out_features + o] = r[j];
        }
}

/* CpuLinearOp::forward dispatch — CpuLinearOp.ixx:248-266 with use_loop_unroll_ = batch % 8 == 0 (:226). */
void oracle_cpu_linear_forward(const float* X, float* Y, const float* W, const float* B,
                               int64_t batch, int64_t in_features, int64_t out_features)
{
    if (batch % 8 == 0) oracle_cpu_linear_forward_unrolled(X, Y, W, B, batch, in_features, out_features);
    else                oracle_cpu_linear_forward_naive(X, Y, W, B, batch, in_features, out_features);
}

/* ------------------------------------------------------------------------- */
/* W4A8 helper restatements (reference's live FP4 prefill numerics)          */
/* ------------------------------------------------------------------------- */

/* sB = max(max_i scale_i, 1e-12) * (6/448) — LIN/Kernels/W4A16Gemm/CudaW4A16Gemm.cu:244-288. */
float oracle_compute_fp8_weight_scale(const float* group_scales, int64_t num_scales)
{
    float m = 0.0f;
    for (int64_t i = 0; i < num_scales; ++i) m = maxf_nan_ignoring(m, group_scales[i]);
    volatile float c = 6.0f / 448.0f;
    return fmaxf(m, 1e-12f) * c;
}

/* W8[n,k] = e4m3( lut[nib] * (s[n,k/g] * (1/sB)) ) — CudaW4A16Gemm.cu:300-326. */
int oracle_fp4_dequantize_to_fp8(const uint8_t* packed, const float* scales, float sB, uint8_t* out,
                                 int64_t N, int64_t K, int group_size)
{
    if (group_size != 64 && group_size != 128) return -1;
    const int64_t G = K / group_size;
    volatile float inv = 1.0f / sB;
    for (int64_t n = 0; n < N; ++n)
        for (int64_t b = 0; b < K / 2; ++b) {
            uint8_t byte = packed[n * (K / 2) + b];
            volatile float s = scales[n * G + (2 * b) / group_size] * inv;
            volatile float lo = oracle_e2m1_to_f32(byte & 0xFu) * s;
            volatile float hi = oracle_e2m1_to_f32((uint8_t)(byte >> 4)) * s;
            out[n * K + 2 * b] = oracle_f32_to_e4m3(lo);
            out[n * K + 2 * b + 1] = oracle_f32_to_e4m3(hi);
        }
    return 0;
}

/* W16[n,k] = bf16_rn( lut[nib] * s[n,k/g] ) — CudaW4A16Gemm.cu:210-235. */
int oracle_fp4_dequantize_to_bf16(const uint8_t* packed, const float* scales, uint16_t* out,
                                  int64_t N, int64_t K, int group_size)
{
    if (group_size != 64 && group_size != 128) return -1;
    const int64_t G = K / group_size;
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) {
            uint8_t byte = packed[n * (K / 2) + k / 2];
            uint8_t nib = (k & 1) ? (uint8_t)(byte >> 4) : (uint8_t)(byte & 0xFu);
            volatile float v = oracle_e2m1_to_f32(nib) * scales[n * G + k / group_size];
            out[n * K + k] = oracle_f32_to_bf16(v);
        }
    return 0;
}

/* W16[n,k] = bf16_rn( f32(W8[n,k]) * s[n] ) — LIN/Kernels/Fp8Prefill/CudaFp8Prefill.cu:64-84. */
void oracle_fp8_dequantize_to_bf16(const uint8_t* w8, const float* scales, uint16_t* out,
                                   int64_t N, int64_t K)
{
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) {
            volatile float v = oracle_e4m3_to_f32(w8[n * K + k]) * scales[n];
            out[n * K + k] = oracle_f32_to_bf16(v);
        }
}

/* Per-token activation quantizer — LIN/Kernels/Fp8Prefill/CudaFp8Prefill.cu:116-162:
 * sA[t] = max(absmax_t, 1e-12)/448 ; X8[t,k] = e4m3( x * (1/sA[t]) ). */
void oracle_quantize_bf16_to_fp8_per_token(const uint16_t* x, uint8_t* x8, float* sA,
                                           int64_t M, int64_t K)
{
    for (int64_t t = 0; t < M; ++t) {
        float absmax = 0.0f;
        for (int64_t k = 0; k < K; ++k)
            absmax = maxf_nan_ignoring(absmax, fabsf(oracle_bf16_to_f32(x[t * K + k])));
        volatile float scale = fmaxf(absmax, 1e-12f) / 448.0f;
        volatile float inv = 1.0f / scale;
        sA[t] = scale;
        for (int64_t k = 0; k < K; ++k) {
            volatile float v = oracle_bf16_to_f32(x[t * K + k]) * inv;
            x8[t * K + k] = oracle_f32_to_e4m3(v);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Gated activations behind the gate|up Linear.  X [tokens, 2H] BF16 with the gate half
 * first, Y [tokens, H] BF16.  FP32 arithmetic on the BF16 inputs, one RN rounding at the store.
 *   GeGLU : Activations/Geglu/Kernels/Geglu.cu:42-61, GeluTanh = ElementwiseActivation.h:41-50
 *   SwiGLU: Activations/Swiglu/Kernels/Swiglu.Bf16.cu:77-83,:165-195 (the device kernel uses __expf and
 *           __frcp_rn; libm expf and an IEEE division differ from them by a few FP32 ulps, so this CPU
 *           restatement is pinned to the reference kernel within one BF16 ulp, not bit for bit).
 * ------------------------------------------------------------------------------------------ */
static float oracle_gelu_tanh(float x)
{
    const float kScale = 0.7978845608f, kCoeff = 0.044715f;
    float cube = kCoeff * x * x * x;
    return 0.5f * x * (1.0f + tanhf(kScale * (x + cube)));
}
static float oracle_silu(float x) { return x * (1.0f / (1.0f + expf(-x))); }

/* kind: 1 = GeGLU (tanh), 2 = SwiGLU */
int oracle_glu_forward_bf16(const uint16_t* X, uint16_t* Y, int64_t tokens, int64_t H, int kind)
{
    if (kind != 1 && kind != 2) return -1;
    for (int64_t t = 0; t < tokens; ++t)
        for (int64_t c = 0; c < H; ++c) {
            const float g = oracle_bf16_to_f32(X[t * 2 * H + c]);
            const float u = oracle_bf16_to_f32(X[t * 2 * H + H + c]);
            const float a = (kind == 1) ? oracle_gelu_tanh(g) : oracle_silu(g);
            Y[t * H + c] = oracle_f32_to_bf16(a * u);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * FP8 tied-table token-embedding gather-dequant:
 * Y[bt,:] = bf16( f32(W8[X[bt],:]) * scales[X[bt]] ) — Embeddings/Kernels/TokenEmbedding.Fp8.cu:34-66.
 * ------------------------------------------------------------------------------------------ */
void oracle_token_embedding_qfp8(const int32_t* X, const uint8_t* w8, const float* scales, uint16_t* Y,
                                 int64_t rows, int64_t C)
{
    for (int64_t bt = 0; bt < rows; ++bt) {
        const int64_t ix = X[bt];
        const float s = scales[ix];
        for (int64_t c = 0; c < C; ++c)
            Y[bt * C + c] = oracle_f32_to_bf16(oracle_e4m3_to_f32(w8[ix * C + c]) * s);
    }
}

/* CONTEXT ROW, NOT THE REFERENCE: the same loop as forwardNaive with a `float` accumulator (what an optimising build
 * of a plain FP32 Linear would do; SURVEY.md 8d "labelled context rows").  bench.py reports it beside the reference
 * figure so that the 80-bit accumulator's cost is visible. */
void oracle_ctx_linear_forward_f32acc(const float* X, float* Y, const float* W, const float* B,
                                      int64_t batch, int64_t in_features, int64_t out_features)
{
    for (int64_t idx = 0; idx < batch; ++idx)
        for (int64_t o = 0; o < out_features; ++o) {
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            const float* x = X + idx * in_features; const float* w = W + o * in_features;
            int64_t i = 0;
            for (; i + 4 <= in_features; i += 4) { a0 += x[i] * w[i]; a1 += x[i + 1] * w[i + 1]; a2 += x[i + 2] * w[i + 2]; a3 += x[i + 3] * w[i + 3]; }
            for (; i < in_features; ++i) a0 += x[i] * w[i];
            Y[idx * out_features + o] = (a0 + a1) + (a2 + a3) + (B ? B[o] : 0.0f);
        }
}

/* PerGroupInt4 W4A16 forward — the arithmetic of fused_w4a16_gemm_kernel<g, false>
 * (Compute/Devices/Cuda/Operations/Linear/Kernels/W4A16Gemm/CudaW4A16Gemm.cu:88-197): per element
 * w = (float(nibble) - zero) * scale rounded to FP32 (:154-177), acc += a * w in k order with the FMA contraction nvcc
 * applies to `acc += a * b` (:189-190), out = bf16(acc + bias) (:196).  zero_points may be NULL (zero = 8). */
void oracle_w4a16_int4_forward(const uint16_t* x_bf16, const uint8_t* w_packed, const float* scales,
                               const uint8_t* zero_points, const uint16_t* bias_bf16, uint16_t* y_bf16, float* y_f32,
                               int64_t M, int64_t K, int64_t N, int group_size)
{
    const int64_t KG = K / group_size;
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            float acc = 0.0f;
            for (int64_t k = 0; k < K; ++k) {
                const uint8_t byte = w_packed[n * (K / 2) + k / 2];
                const float nib = (float)((k & 1) ? (byte >> 4) : (byte & 0xF));
                const int64_t grp = k / group_size;
                float zero = 8.0f;
                if (zero_points) {
                    const uint8_t zb = zero_points[n * (KG / 2) + grp / 2];
                    zero = (float)((grp & 1) ? (zb >> 4) : (zb & 0xF));
                }
                const float w = (nib - zero) * scales[n * KG + grp];
                acc = fmaf(oracle_bf16_to_f32(x_bf16[m * K + k]), w, acc);
            }
            const float r = acc + (bias_bf16 ? oracle_bf16_to_f32(bias_bf16[n]) : 0.0f);
            if (y_f32) y_f32[m * N + n] = r;
            y_bf16[m * N + n] = oracle_f32_to_bf16(r);
        }
}

/* BF16 RMSNorm over rows of length K — Normalizations/RmsNorm/Kernels/RmsNorm.Bf16.cu:19-73 restated: lane-strided FMA
 * partial sums (lane i: x[i], x[i+32], ...), the shfl_down tree 16/8/4/2/1 of lane 0, rstd = 1/sqrt(m2/K + eps),
 * out = bf16(fma(x * rstd, w + offset, b)).  The device uses MUFU.RSQ (rsqrtf, <= 2 ulp) where this uses 1/sqrtf, so the
 * restatement is pinned to the reference kernel's golden outputs within one BF16 ulp, not bit for bit
 * (tests/golden/rmsnorm_*.npz; the GPU tests compare the CUDA kernels with the reference kernel itself bit for bit). */
void oracle_rmsnorm_forward_bf16(const uint16_t* x, const uint16_t* weight, const uint16_t* bias, uint16_t* y, float* rstd_out,
                                 int64_t M, int64_t K, float eps, float weight_offset)
{
    for (int64_t m = 0; m < M; ++m) {
        float part[32];
        for (int lane = 0; lane < 32; ++lane) {
            float a = 0.0f;
            for (int64_t i = lane; i < K; i += 32) { const float v = oracle_bf16_to_f32(x[m * K + i]); a = fmaf(v, v, a); }
            part[lane] = a;
        }
        for (int off = 16; off > 0; off >>= 1)
            for (int lane = 0; lane < off; ++lane) part[lane] = part[lane] + part[lane + off];   /* lane 0's tree */
        const float rstd = 1.0f / sqrtf(part[0] / (float)K + eps);
        if (rstd_out) rstd_out[m] = rstd;
        for (int64_t i = 0; i < K; ++i) {
            const float w = weight ? oracle_bf16_to_f32(weight[i]) + weight_offset : 1.0f;
            const float b = bias ? oracle_bf16_to_f32(bias[i]) : 0.0f;
            const float t = oracle_bf16_to_f32(x[m * K + i]) * rstd;
            y[m * K + i] = oracle_f32_to_bf16(fmaf(t, w, b));
        }
    }
}
