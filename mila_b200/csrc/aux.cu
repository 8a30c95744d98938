// aux.cu — staging / W4A8 helper kernels that Mila's 2-phase paths call around cuBLASLt.
// They are not on our fused hot path (the fused GEMV/GEMM never materialises dequantised
// weights) but CudaLinearOp.ixx references the symbols, so they are provided, bit-exact with:
//   cuda_fp8_dequantize_to_bf16           LIN/Kernels/Fp8Prefill/CudaFp8Prefill.cu:64-100
//   cuda_fp4_dequantize_to_bf16           LIN/Kernels/W4A16Gemm/CudaW4A16Gemm.cu:210-235
//   cuda_compute_fp8_weight_scale         CudaW4A16Gemm.cu:244-288
//   cuda_fp4_dequantize_to_fp8            CudaW4A16Gemm.cu:300-326
//   cuda_quantize_bf16_to_fp8_per_token   CudaFp8Prefill.cu:116-162
//   cuda_fp8_apply_per_token_scales       CudaFp8Prefill.cu:191-211
//   cuda_add_bias (bf16)                  CudaFp8Prefill.cu:239-256
// All are flat, vectorised streaming kernels (HBM-bound).
#include "common.cuh"

namespace milab200 {
namespace {

__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ uint16_t e4m3x2_rn_sat(float lo, float hi)
{
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(r) : "f"(lo), "f"(hi));
    return r;
}

__device__ __forceinline__ float e4m3_to_f32(uint32_t byte)
{
    uint32_t h2;
    asm("{ .reg .b16 l; cvt.u16.u32 l, %1; cvt.rn.f16x2.e4m3x2 %0, l; }" : "=r"(h2) : "r"(byte));
    return __half2float(__ushort_as_half((unsigned short)(h2 & 0xFFFFu)));
}

// out[n, k] = bf16( f32(w8[n,k]) * s[n] ); 8 elements per thread.
__global__ void __launch_bounds__(256)
fp8_dequant_bf16_kernel(uint4* __restrict__ out, const uint2* __restrict__ w8, const float* __restrict__ scales,
                        int64_t chunks, int chunks_per_row)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = w8[c];
        const float s = __ldg(scales + c / chunks_per_row);
        uint32_t a, b, d, e;
        cvt_e4m3x4_to_f16x2x2(v.x, a, b);
        cvt_e4m3x4_to_f16x2x2(v.y, d, e);
        const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&a));
        const float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&b));
        const float2 f2 = __half22float2(*reinterpret_cast<__half2*>(&d));
        const float2 f3 = __half22float2(*reinterpret_cast<__half2*>(&e));
        uint4 o;
        o.x = pack_bf16x2_rn(f0.x * s, f0.y * s);
        o.y = pack_bf16x2_rn(f1.x * s, f1.y * s);
        o.z = pack_bf16x2_rn(f2.x * s, f2.y * s);
        o.w = pack_bf16x2_rn(f3.x * s, f3.y * s);
        out[c] = o;
    }
}

// 8 nibbles (one word) per thread.  MODE 0: bf16 out, MODE 1: e4m3 out with s * (1/sB).
template <int MODE>
__global__ void __launch_bounds__(256)
fp4_dequant_kernel(void* __restrict__ out, const uint32_t* __restrict__ packed, const float* __restrict__ scales,
                   const float* __restrict__ sB, int64_t words, int words_per_group)
{
    float inv = 1.0f;
    if (MODE == 1) inv = __fdiv_rn(1.0f, *sB);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < words; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = packed[c];
        float s = __ldg(scales + c / words_per_group);
        if (MODE == 1) s = s * inv;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = e2m1_to_f32((w >> (4 * i)) & 0xFu) * s;
        if (MODE == 0) {
            uint4 o;
            o.x = pack_bf16x2_rn(f[0], f[1]); o.y = pack_bf16x2_rn(f[2], f[3]);
            o.z = pack_bf16x2_rn(f[4], f[5]); o.w = pack_bf16x2_rn(f[6], f[7]);
            reinterpret_cast<uint4*>(out)[c] = o;
        } else {
            uint2 o;
            o.x = (uint32_t)e4m3x2_rn_sat(f[0], f[1]) | ((uint32_t)e4m3x2_rn_sat(f[2], f[3]) << 16);
            o.y = (uint32_t)e4m3x2_rn_sat(f[4], f[5]) | ((uint32_t)e4m3x2_rn_sat(f[6], f[7]) << 16);
            reinterpret_cast<uint2*>(out)[c] = o;
        }
    }
}

// max over non-negative scales via integer atomicMax; `out` must be zeroed first.
__global__ void __launch_bounds__(256)
scale_max_kernel(const float* __restrict__ s, int64_t n, float* __restrict__ out)
{
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, s[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
__global__ void scale_finalize_kernel(float* out)
{
    *out = fmaxf(*out, 1e-12f) * (6.0f / 448.0f);
}

// one CTA per token: absmax -> sA = max(absmax,1e-12)/448 -> x8 = e4m3(x * (1/sA))
__global__ void __launch_bounds__(256)
act_quant_per_token_kernel(uint8_t* __restrict__ x8, float* __restrict__ sA, const __nv_bfloat16* __restrict__ x, int K)
{
    __shared__ float s_red[8];
    const int64_t row = blockIdx.x;
    const __nv_bfloat16* xr = x + row * K;
    uint8_t* orow = x8 + row * K;
    const int tid = threadIdx.x;
    const bool vec = (K % 8 == 0);
    float m = 0.0f;
    if (vec) {
        for (int c = tid; c < K / 8; c += 256) {
            const uint4 v = ldg_cached_v4(xr + c * 8);
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int i = 0; i < 4; ++i) { m = fmaxf(m, fabsf(bf16lo(w[i]))); m = fmaxf(m, fabsf(bf16hi(w[i]))); }
        }
    } else {
        for (int k = tid; k < K; k += 256) m = fmaxf(m, fabsf(__bfloat162float(xr[k])));
    }
    m = warp_max(m);
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    __syncthreads();
    float absmax = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) absmax = fmaxf(absmax, s_red[w]);
    const float scale = __fdiv_rn(fmaxf(absmax, 1e-12f), 448.0f);
    const float inv = __fdiv_rn(1.0f, scale);
    if (tid == 0) sA[row] = scale;
    if (vec) {
        for (int c = tid; c < K / 8; c += 256) {
            const uint4 v = ldg_cached_v4(xr + c * 8);
            uint2 o;
            o.x = (uint32_t)e4m3x2_rn_sat(bf16lo(v.x) * inv, bf16hi(v.x) * inv) |
                  ((uint32_t)e4m3x2_rn_sat(bf16lo(v.y) * inv, bf16hi(v.y) * inv) << 16);
            o.y = (uint32_t)e4m3x2_rn_sat(bf16lo(v.z) * inv, bf16hi(v.z) * inv) |
                  ((uint32_t)e4m3x2_rn_sat(bf16lo(v.w) * inv, bf16hi(v.w) * inv) << 16);
            *reinterpret_cast<uint2*>(orow + c * 8) = o;
        }
    } else {
        for (int k = tid; k < K; k += 256)
            orow[k] = (uint8_t)(e4m3x2_rn_sat(__bfloat162float(xr[k]) * inv, 0.0f) & 0xFFu);
    }
}

// y[t,n] = bf16( f32(y[t,n]) * sA[t] (+ bias[n]) )   /   y[t,n] = bf16( f32(y) + f32(bias) )
template <bool kScale>
__global__ void __launch_bounds__(256)
row_epilogue_kernel(__nv_bfloat16* __restrict__ y, const float* __restrict__ sA,
                    const __nv_bfloat16* __restrict__ bias, int N)
{
    __nv_bfloat16* row = y + (int64_t)blockIdx.y * N;
    const float s = kScale ? sA[blockIdx.y] : 1.0f;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float v = __bfloat162float(row[n]);
        if (kScale) v = v * s;
        if (bias) v += __bfloat162float(bias[n]);
        row[n] = __float2bfloat16_rn(v);
    }
}

int stream_grid(int64_t items)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t need = (items + 255) / 256;
    const int64_t cap = (int64_t)sms * 8;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace

int launch_fp8_dequantize_to_bf16(void* out, const void* w8, const float* scales, int N, int K, cudaStream_t st)
{
    if (!out || !w8 || !scales || N <= 0 || K <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (K % 8 != 0) return MILAB200_E_BAD_SHAPE;
    const int64_t chunks = (int64_t)N * K / 8;
    fp8_dequant_bf16_kernel<<<stream_grid(chunks), 256, 0, st>>>(
        static_cast<uint4*>(out), static_cast<const uint2*>(w8), scales, chunks, K / 8);
    note_launch("fp8_dequant_bf16_kernel");
    return (int)cudaGetLastError();
}

int launch_fp4_dequantize_to_bf16(void* out, const void* packed, const float* scales, int N, int K, int g, cudaStream_t st)
{
    if (!out || !packed || !scales || N <= 0 || K <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (g != 64 && g != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    if (K % g != 0) return MILAB200_E_BAD_SHAPE;
    const int64_t words = (int64_t)N * K / 8;
    fp4_dequant_kernel<0><<<stream_grid(words), 256, 0, st>>>(out, static_cast<const uint32_t*>(packed), scales, nullptr, words, g / 8);
    note_launch("fp4_dequant_kernel<bf16>");
    return (int)cudaGetLastError();
}

int launch_compute_fp8_weight_scale(float* out, const float* group_scales, int64_t n, cudaStream_t st)
{
    if (!out || !group_scales || n <= 0) return MILAB200_E_INVALID_ARGUMENT;
    MILAB200_RETURN_IF_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    int grid = stream_grid(n); if (grid > 1024) grid = 1024;
    scale_max_kernel<<<grid, 256, 0, st>>>(group_scales, n, out);
    scale_finalize_kernel<<<1, 1, 0, st>>>(out);
    note_launch("scale_max_kernel+finalize", 2);
    return (int)cudaGetLastError();
}

int launch_fp4_dequantize_to_fp8(void* out, const void* packed, const float* scales, const float* sB,
                                 int N, int K, int g, cudaStream_t st)
{
    if (!out || !packed || !scales || !sB || N <= 0 || K <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (g != 64 && g != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    if (K % g != 0) return MILAB200_E_BAD_SHAPE;
    const int64_t words = (int64_t)N * K / 8;
    fp4_dequant_kernel<1><<<stream_grid(words), 256, 0, st>>>(out, static_cast<const uint32_t*>(packed), scales, sB, words, g / 8);
    note_launch("fp4_dequant_kernel<fp8>");
    return (int)cudaGetLastError();
}

int launch_quantize_bf16_to_fp8_per_token(void* x8, float* sA, const void* x, int M, int K, cudaStream_t st)
{
    if (!x8 || !sA || !x || M <= 0 || K <= 0) return MILAB200_E_INVALID_ARGUMENT;
    act_quant_per_token_kernel<<<M, 256, 0, st>>>(static_cast<uint8_t*>(x8), sA, static_cast<const __nv_bfloat16*>(x), K);
    note_launch("act_quant_per_token_kernel");
    return (int)cudaGetLastError();
}

int launch_fp8_apply_per_token_scales(void* y, const float* sA, const void* bias, int M, int N, cudaStream_t st)
{
    if (!y || !sA || M <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    dim3 grid((N + 255) / 256 > 64 ? 64 : (N + 255) / 256, M);
    row_epilogue_kernel<true><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(y), sA, static_cast<const __nv_bfloat16*>(bias), N);
    note_launch("row_epilogue_kernel<scale>");
    return (int)cudaGetLastError();
}

// FP32 overload of cuda_add_bias (CudaFp8Prefill.cuh:151; the non-quantized FP32 prefill path adds its bias post-GEMM)
__global__ void __launch_bounds__(256)
add_bias_f32_kernel(float* __restrict__ y, const float* __restrict__ bias, int N)
{
    float* row = y + (int64_t)blockIdx.y * N;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) row[n] = row[n] + bias[n];
}

int launch_add_bias_f32(float* y, const float* bias, int M, int N, cudaStream_t st)
{
    if (!y || !bias || M <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    dim3 grid((N + 255) / 256 > 64 ? 64 : (N + 255) / 256, M);
    add_bias_f32_kernel<<<grid, 256, 0, st>>>(y, bias, N);
    note_launch("add_bias_f32_kernel");
    return (int)cudaGetLastError();
}

int launch_add_bias_bf16(void* y, const void* bias, int M, int N, cudaStream_t st)
{
    if (!y || !bias || M <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    dim3 grid((N + 255) / 256 > 64 ? 64 : (N + 255) / 256, M);
    row_epilogue_kernel<false><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(y), nullptr, static_cast<const __nv_bfloat16*>(bias), N);
    note_launch("row_epilogue_kernel<bias>");
    return (int)cudaGetLastError();
}

}  // namespace milab200
