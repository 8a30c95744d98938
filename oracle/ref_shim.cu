// ref_shim.cu — extern "C" doorway into the REFERENCE's own CUDA launchers, compiled
// unmodified from /root/reference by oracle/Makefile (`make ref`) into
// oracle/_ref/libmila_ref_linear.so.  Test infrastructure only: it lets the GPU parity
// tests compare our packed bytes / scales / outputs with what Mila's kernels produce on
// the same inputs, and lets bench.py time the recompiled reference kernels as context.
// This file contains no reference code — only forwarding calls; the headers it includes
// are read from the reference tree at build time (-I), never copied.
#include <cstdint>
#include <cstdio>
#include <exception>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include "Quantization/CudaFp8WeightQuantization.cuh"
#include "Quantization/CudaFp4WeightQuantization.cuh"
#include "W8A16Gemm/CudaW8A16Gemm.cuh"
#include "W4A16Gemm/CudaW4A16Gemm.cuh"
#include "W4A16Gemm/CudaW4A16Gemm.Wmma.cuh"
#include "Fp8Prefill/CudaFp8Prefill.cuh"

// Linear.cuh drags in <cublasLt.h>; declare the two matvecs we need instead of including it.
namespace Mila::Dnn::Compute::Cuda::Linear {
    void cuda_matvec_decode_bf16_qfp8(__nv_bfloat16*, const __nv_bfloat16*, const __nv_fp8_e4m3*,
                                      const float*, const __nv_bfloat16*, int, int, cudaStream_t);
    void cuda_matvec_decode_bf16_qfp4(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*,
                                      const float*, const __nv_bfloat16*, int, int, int, cudaStream_t);
}

// gated activations behind the gate|up Linear (Activations/{Geglu,Swiglu}/Kernels): declared, not included —
// their headers sit in other include roots
namespace Mila::Dnn::Compute::Cuda::Geglu {
    void cuda_geglu_forward_bf16(__nv_bfloat16* Y, const __nv_bfloat16* X, int N, int half_width, cudaStream_t stream);
}
namespace Mila::Dnn::Compute::Cuda::Swiglu {
    void cuda_swiglu_forward_bf16(__nv_bfloat16* Y, const __nv_bfloat16* X, int N, int half_width, cudaStream_t stream);
}

namespace Mila::Dnn::Compute::Cuda::TokenEmbedding {
    void cuda_token_embedding_forward_bf16_qfp8(__nv_bfloat16* Y, const int* X, const void* wte_fp8, const float* scales,
                                                int B, int T, int C, cudaStream_t stream);
}

namespace Mila::Dnn::Compute::Cuda::RmsNorm {
    void cuda_rmsnorm_forward_bf16(__nv_bfloat16* Y, __nv_bfloat16* rstd, const __nv_bfloat16* X, const __nv_bfloat16* weight,
                                   const __nv_bfloat16* bias, int outer_size, int inner_size, int norm_dim, float epsilon,
                                   float weight_offset, cudaStream_t stream);
}

namespace ref = Mila::Dnn::Compute::Cuda::Linear;

#define REF_GUARD(stmt)                                                        \
    try { stmt; } catch (const std::exception& e) {                            \
        std::fprintf(stderr, "[mila_ref] %s\n", e.what()); return -1; }        \
    return (int)cudaGetLastError();

extern "C" {

int milaref_quantize_fp8_per_channel(const void* src_host, void* dst, float* scales,
        int64_t N, int64_t K, void* staging, void* stream)
{ REF_GUARD(ref::cuda_quantize_fp8_per_channel(src_host, dst, scales, N, K, staging, (cudaStream_t)stream)) }

int milaref_quantize_fp4_per_group(const void* src_host, void* dst, float* scales,
        int64_t N, int64_t K, int g, void* staging, void* stream)
{ REF_GUARD(ref::cuda_quantize_fp4_per_group(src_host, dst, scales, N, K, g, staging, (cudaStream_t)stream)) }

int milaref_matvec_decode_bf16_qfp8(void* y, const void* x, const void* W, const float* scales,
        const void* bias, int C, int OC, void* stream)
{ REF_GUARD(ref::cuda_matvec_decode_bf16_qfp8((__nv_bfloat16*)y, (const __nv_bfloat16*)x,
        (const __nv_fp8_e4m3*)W, scales, (const __nv_bfloat16*)bias, C, OC, (cudaStream_t)stream)) }

int milaref_matvec_decode_bf16_qfp4(void* y, const void* x, const void* W, const float* scales,
        const void* bias, int C, int OC, int g, void* stream)
{ REF_GUARD(ref::cuda_matvec_decode_bf16_qfp4((__nv_bfloat16*)y, (const __nv_bfloat16*)x,
        (const uint8_t*)W, scales, (const __nv_bfloat16*)bias, C, OC, g, (cudaStream_t)stream)) }

int milaref_w8a16_gemm(void* out, const void* act, const void* W, const float* scales,
        const void* bias, int M, int K, int N, void* stream)
{ REF_GUARD(ref::cuda_w8a16_gemm((__nv_bfloat16*)out, (const __nv_bfloat16*)act,
        (const __nv_fp8_e4m3*)W, scales, (const __nv_bfloat16*)bias, M, K, N, (cudaStream_t)stream)) }

int milaref_fp4a16_gemm(void* out, const void* act, const void* W, const float* scales,
        const void* bias, int M, int K, int N, int g, void* stream)
{ REF_GUARD(ref::cuda_fp4a16_gemm((__nv_bfloat16*)out, (const __nv_bfloat16*)act,
        (const uint8_t*)W, scales, (const __nv_bfloat16*)bias, M, K, N, g, (cudaStream_t)stream)) }

int milaref_w4a16_gemm(void* out, const void* act, const void* W, const float* scales, const void* zero_points,
        const void* bias, int M, int K, int N, int g, void* stream)
{ REF_GUARD(ref::cuda_w4a16_gemm((__nv_bfloat16*)out, (const __nv_bfloat16*)act, (const uint8_t*)W, scales,
        (const uint8_t*)zero_points, (const __nv_bfloat16*)bias, M, K, N, g, (cudaStream_t)stream)) }

int milaref_fp4a16_gemm_wmma(void* out, const void* act, const void* W, const float* scales,
        const void* bias, int M, int K, int N, int g, void* stream)
{ REF_GUARD(ref::cuda_fp4a16_gemm_wmma((__nv_bfloat16*)out, (const __nv_bfloat16*)act,
        (const uint8_t*)W, scales, (const __nv_bfloat16*)bias, M, K, N, g, (cudaStream_t)stream)) }

int milaref_fp8_dequantize_to_bf16(void* out, const void* W, const float* scales, int N, int K, void* stream)
{ REF_GUARD(ref::cuda_fp8_dequantize_to_bf16((__nv_bfloat16*)out, (const __nv_fp8_e4m3*)W, scales, N, K, (cudaStream_t)stream)) }

int milaref_fp4_dequantize_to_bf16(void* out, const void* W, const float* scales, int N, int K, int g, void* stream)
{ REF_GUARD(ref::cuda_fp4_dequantize_to_bf16((__nv_bfloat16*)out, (const uint8_t*)W, scales, N, K, g, (cudaStream_t)stream)) }

int milaref_compute_fp8_weight_scale(float* out, const float* group_scales, int64_t n, void* stream)
{ REF_GUARD(ref::cuda_compute_fp8_weight_scale(out, group_scales, n, (cudaStream_t)stream)) }

int milaref_fp4_dequantize_to_fp8(void* out, const void* W, const float* scales, const float* sB,
        int N, int K, int g, void* stream)
{ REF_GUARD(ref::cuda_fp4_dequantize_to_fp8((__nv_fp8_e4m3*)out, (const uint8_t*)W, scales, sB, N, K, g, (cudaStream_t)stream)) }

int milaref_quantize_bf16_to_fp8_per_token(void* x8, float* sA, const void* x, int M, int K, void* stream)
{ REF_GUARD(ref::cuda_quantize_bf16_to_fp8_per_token((__nv_fp8_e4m3*)x8, sA, (const __nv_bfloat16*)x, M, K, (cudaStream_t)stream)) }

int milaref_fp8_apply_per_token_scales(void* out, const float* sA, const void* bias, int M, int N, void* stream)
{ REF_GUARD(ref::cuda_fp8_apply_per_token_scales((__nv_bfloat16*)out, sA, (const __nv_bfloat16*)bias, M, N, (cudaStream_t)stream)) }


int milaref_token_embedding_forward_bf16_qfp8(void* Y, const int* X, const void* wte, const float* scales,
        int B, int T, int C, void* stream)
{ REF_GUARD(Mila::Dnn::Compute::Cuda::TokenEmbedding::cuda_token_embedding_forward_bf16_qfp8((__nv_bfloat16*)Y, X, wte,
        scales, B, T, C, (cudaStream_t)stream)) }

int milaref_geglu_forward_bf16(void* Y, const void* X, int N, int half_width, void* stream)
{ REF_GUARD(Mila::Dnn::Compute::Cuda::Geglu::cuda_geglu_forward_bf16((__nv_bfloat16*)Y, (const __nv_bfloat16*)X,
        N, half_width, (cudaStream_t)stream)) }

int milaref_swiglu_forward_bf16(void* Y, const void* X, int N, int half_width, void* stream)
{ REF_GUARD(Mila::Dnn::Compute::Cuda::Swiglu::cuda_swiglu_forward_bf16((__nv_bfloat16*)Y, (const __nv_bfloat16*)X,
        N, half_width, (cudaStream_t)stream)) }

int milaref_rmsnorm_forward_bf16(void* Y, void* rstd, const void* X, const void* weight, const void* bias,
        int outer, int inner, int norm_dim, float eps, float weight_offset, void* stream)
{ REF_GUARD(Mila::Dnn::Compute::Cuda::RmsNorm::cuda_rmsnorm_forward_bf16((__nv_bfloat16*)Y, (__nv_bfloat16*)rstd,
        (const __nv_bfloat16*)X, (const __nv_bfloat16*)weight, (const __nv_bfloat16*)bias, outer, inner, norm_dim, eps,
        weight_offset, (cudaStream_t)stream)) }

}  // extern "C"
