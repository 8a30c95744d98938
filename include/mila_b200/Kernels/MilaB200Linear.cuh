/*
 * MilaB200Linear.cuh — source-compatible drop-in for Mila's kernel-launcher headers.
 *
 * Declares, in Mila's own namespace and with Mila's exact names and signatures, the launchers that
 * CudaLinearOp.ixx / CudaTokenEmbeddingOp.Quantize.ixx call, as inline forwards to the C-ABI of
 * libmila_b200_linear.so.  A maintainer replaces
 *     #include "Kernels/Linear.cuh"                       (K/Linear.cuh:49-81)
 *     #include "Kernels/Quantization/CudaFp8WeightQuantization.cuh"   (:56-63)
 *     #include "Kernels/Quantization/CudaFp4WeightQuantization.cuh"   (:52-60)
 *     #include "Kernels/W8A16Gemm/CudaW8A16Gemm.cuh"      (:59-68)
 *     #include "Kernels/W4A16Gemm/CudaW4A16Gemm.cuh"      (:119-129 and the helper launchers)
 *     #include "Kernels/W4A16Gemm/CudaW4A16Gemm.Wmma.cuh" (:46-56)
 *     #include "Kernels/Fp8Prefill/CudaFp8Prefill.cuh"
 * with this header and links -lmila_b200_linear; see INTEGRATION.md.
 *
 * Error convention preserved: the reference launchers `throw std::runtime_error` on CUDA failures and
 * unsupported group sizes (CudaFp8WeightQuantization.cu:223-248, CudaFp4WeightQuantization.cu:220);
 * a non-zero C-ABI return becomes exactly that.  Unlike the reference's GEMM launchers, an
 * unsupported group_size is never silently ignored (CudaW4A16Gemm.cu:361-363).
 */
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include "../../mila_b200_linear.h"

namespace Mila::Dnn::Compute::Cuda::Linear
{
    namespace milab200_detail
    {
        inline void check( int rc, const char* what )
        {
            if ( rc != 0 )
                throw std::runtime_error( std::string( what ) + ": " + milab200_error_string( rc ) );
        }
    }

    // ---- load-time quantizers -------------------------------------------------------------
    inline void cuda_quantize_fp8_per_channel(
        const void* src_bf16, void* dst_fp8, float* dst_scales,
        int64_t out_features, int64_t in_features, void* dev_staging, cudaStream_t stream )
    {
        milab200_detail::check( milab200_quantize_fp8_per_channel(
            src_bf16, dst_fp8, dst_scales, out_features, in_features, dev_staging, stream ),
            "cuda_quantize_fp8_per_channel" );
    }

    inline void cuda_quantize_fp4_per_group(
        const void* src_bf16, void* dst_packed, float* dst_scales,
        int64_t out_features, int64_t in_features, int group_size, void* dev_staging, cudaStream_t stream )
    {
        milab200_detail::check( milab200_quantize_fp4_per_group(
            src_bf16, dst_packed, dst_scales, out_features, in_features, group_size, dev_staging, stream ),
            "cuda_quantize_fp4_per_group" );
    }

    // ---- decode matvecs -------------------------------------------------------------------
    inline void cuda_matvec_decode_bf16_qfp8(
        __nv_bfloat16* y, const __nv_bfloat16* x, const __nv_fp8_e4m3* weight, const float* scales,
        const __nv_bfloat16* bias, int C, int OC, cudaStream_t stream )
    {
        milab200_detail::check( milab200_matvec_decode_bf16_qfp8( y, x, weight, scales, bias, C, OC, stream ),
            "cuda_matvec_decode_bf16_qfp8" );
    }

    inline void cuda_matvec_decode_bf16_qfp4(
        __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* weights_packed, const float* scales,
        const __nv_bfloat16* bias, int C, int OC, int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_matvec_decode_bf16_qfp4(
            y, x, weights_packed, scales, bias, C, OC, group_size, stream ),
            "cuda_matvec_decode_bf16_qfp4" );
    }

    // ---- batched fused slots (kUseW8A16Gemm / kUseFusedFp4Gemm = true) ------------------------
    inline void cuda_w8a16_gemm(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const __nv_fp8_e4m3* weights,
        const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int out_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_w8a16_gemm(
            output, activations, weights, scales, bias, outer_size, in_features, out_features, stream ),
            "cuda_w8a16_gemm" );
    }

    inline void cuda_fp4a16_gemm(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const uint8_t* weights_packed,
        const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int out_features, int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4a16_gemm(
            output, activations, weights_packed, scales, bias,
            outer_size, in_features, out_features, group_size, stream ), "cuda_fp4a16_gemm" );
    }

    inline void cuda_fp4a16_gemm_wmma(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const uint8_t* weights_packed,
        const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int out_features, int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4a16_gemm_wmma(
            output, activations, weights_packed, scales, bias,
            outer_size, in_features, out_features, group_size, stream ), "cuda_fp4a16_gemm_wmma" );
    }

    // ---- gate|up Linear + gated activation in one launch (new surface) ----------
    inline void cuda_w8a16_gemm_glu(
        __nv_bfloat16* output, __nv_bfloat16* gate_up_scratch, const __nv_bfloat16* activations,
        const __nv_fp8_e4m3* weights, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int hidden, int glu_kind, cudaStream_t stream )
    {
        milab200_detail::check( milab200_w8a16_gemm_glu( output, gate_up_scratch, activations, weights, scales, bias,
            outer_size, in_features, hidden, glu_kind, stream ), "cuda_w8a16_gemm_glu" );
    }

    inline void cuda_fp4a16_gemm_glu(
        __nv_bfloat16* output, __nv_bfloat16* gate_up_scratch, const __nv_bfloat16* activations,
        const uint8_t* weights_packed, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int hidden, int group_size, int glu_kind, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4a16_gemm_glu( output, gate_up_scratch, activations, weights_packed, scales,
            bias, outer_size, in_features, hidden, group_size, glu_kind, stream ), "cuda_fp4a16_gemm_glu" );
    }

    // ---- tensor-parallel row-parallel slots (new surface; the reference is single-GPU) -----------
    // Same arguments as the batched slots plus the opaque context of milab200_tp_create(); M <= 16.
    inline void cuda_w8a16_gemm_rowparallel(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const __nv_fp8_e4m3* weights_shard,
        const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features_local, int out_features, void* tp_ctx, cudaStream_t stream )
    {
        milab200_detail::check( milab200_w8a16_gemm_rowparallel(
            output, activations, weights_shard, scales, bias, outer_size, in_features_local, out_features,
            tp_ctx, stream ), "cuda_w8a16_gemm_rowparallel" );
    }

    inline void cuda_fp4a16_gemm_rowparallel(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const uint8_t* weights_packed_shard,
        const float* scales_shard, const __nv_bfloat16* bias,
        int outer_size, int in_features_local, int out_features, int group_size, void* tp_ctx,
        cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4a16_gemm_rowparallel(
            output, activations, weights_packed_shard, scales_shard, bias,
            outer_size, in_features_local, out_features, group_size, tp_ctx, stream ),
            "cuda_fp4a16_gemm_rowparallel" );
    }

    // ---- 2-phase staging helpers (only reached if the reference toggles keep those paths) ------
    inline void cuda_fp8_dequantize_to_bf16(
        __nv_bfloat16* output, const __nv_fp8_e4m3* input, const float* scales,
        int out_features, int in_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp8_dequantize_to_bf16(
            output, input, scales, out_features, in_features, stream ), "cuda_fp8_dequantize_to_bf16" );
    }

    inline void cuda_fp4_dequantize_to_bf16(
        __nv_bfloat16* output, const uint8_t* weights_packed, const float* scales,
        int out_features, int in_features, int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4_dequantize_to_bf16(
            output, weights_packed, scales, out_features, in_features, group_size, stream ),
            "cuda_fp4_dequantize_to_bf16" );
    }

    inline void cuda_compute_fp8_weight_scale(
        float* weight_fp8_scale_out, const float* fp4_group_scales, int64_t num_scales, cudaStream_t stream )
    {
        milab200_detail::check( milab200_compute_fp8_weight_scale(
            weight_fp8_scale_out, fp4_group_scales, num_scales, stream ), "cuda_compute_fp8_weight_scale" );
    }

    inline void cuda_fp4_dequantize_to_fp8(
        __nv_fp8_e4m3* output, const uint8_t* weights_packed, const float* scales,
        const float* weight_fp8_scale, int out_features, int in_features, int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4_dequantize_to_fp8(
            output, weights_packed, scales, weight_fp8_scale, out_features, in_features, group_size, stream ),
            "cuda_fp4_dequantize_to_fp8" );
    }

    inline void cuda_quantize_bf16_to_fp8_per_token(
        __nv_fp8_e4m3* fp8_out, float* scales_out, const __nv_bfloat16* input,
        int outer_size, int in_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_quantize_bf16_to_fp8_per_token(
            fp8_out, scales_out, input, outer_size, in_features, stream ),
            "cuda_quantize_bf16_to_fp8_per_token" );
    }

    inline void cuda_fp8_apply_per_token_scales(
        __nv_bfloat16* output, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int out_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp8_apply_per_token_scales(
            output, scales, bias, outer_size, out_features, stream ), "cuda_fp8_apply_per_token_scales" );
    }

    inline void cuda_add_bias(
        __nv_bfloat16* output, const __nv_bfloat16* bias, int outer_size, int out_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_add_bias_bf16( output, bias, outer_size, out_features, stream ),
            "cuda_add_bias" );
    }

    // FP32 overload (K/Fp8Prefill/CudaFp8Prefill.cuh:151)
    inline void cuda_add_bias(
        float* output, const float* bias, int outer_size, int out_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_add_bias_f32( output, bias, outer_size, out_features, stream ), "cuda_add_bias" );
    }

    // ---- PerGroupInt4 (K/W4A16Gemm/CudaW4A16Gemm.cuh:73) --------------------------------------------------
    inline void cuda_w4a16_gemm(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const uint8_t* weights_packed, const float* scales,
        const uint8_t* zero_points, const __nv_bfloat16* bias, int outer_size, int in_features, int out_features,
        int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_w4a16_gemm( output, activations, weights_packed, scales, zero_points, bias,
            outer_size, in_features, out_features, group_size, stream ), "cuda_w4a16_gemm" );
    }

    // ---- batched row-parallel shard + NCCL all-reduce (new surface) --------------------------
    inline void cuda_w8a16_gemm_rowparallel_nccl(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const __nv_fp8_e4m3* weight_shard, const float* scales,
        const __nv_bfloat16* bias, int outer_size, int in_features_local, int out_features, void* nccl_comm, cudaStream_t stream )
    {
        milab200_detail::check( milab200_w8a16_gemm_rowparallel_nccl( output, activations, weight_shard, scales, bias,
            outer_size, in_features_local, out_features, nccl_comm, stream ), "cuda_w8a16_gemm_rowparallel_nccl" );
    }
    inline void cuda_fp4a16_gemm_rowparallel_nccl(
        __nv_bfloat16* output, const __nv_bfloat16* activations, const uint8_t* weights_packed_shard, const float* scales_shard,
        const __nv_bfloat16* bias, int outer_size, int in_features_local, int out_features, int group_size, void* nccl_comm,
        cudaStream_t stream )
    {
        milab200_detail::check( milab200_fp4a16_gemm_rowparallel_nccl( output, activations, weights_packed_shard, scales_shard,
            bias, outer_size, in_features_local, out_features, group_size, nccl_comm, stream ), "cuda_fp4a16_gemm_rowparallel_nccl" );
    }

    // ---- RMSNorm -> Linear as one call -------------------------------------------------
    inline void cuda_rmsnorm_w8a16_gemm(
        __nv_bfloat16* output, __nv_bfloat16* normed_scratch, const __nv_bfloat16* activations,
        const __nv_bfloat16* norm_weight, const __nv_bfloat16* norm_bias, float epsilon, float weight_offset,
        const __nv_fp8_e4m3* weight, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int out_features, cudaStream_t stream )
    {
        milab200_detail::check( milab200_rmsnorm_w8a16_gemm( output, normed_scratch, activations, norm_weight, norm_bias, epsilon,
            weight_offset, weight, scales, bias, outer_size, in_features, out_features, stream ), "cuda_rmsnorm_w8a16_gemm" );
    }
    inline void cuda_rmsnorm_fp4a16_gemm(
        __nv_bfloat16* output, __nv_bfloat16* normed_scratch, const __nv_bfloat16* activations,
        const __nv_bfloat16* norm_weight, const __nv_bfloat16* norm_bias, float epsilon, float weight_offset,
        const uint8_t* weights_packed, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int out_features, int group_size, cudaStream_t stream )
    {
        milab200_detail::check( milab200_rmsnorm_fp4a16_gemm( output, normed_scratch, activations, norm_weight, norm_bias, epsilon,
            weight_offset, weights_packed, scales, bias, outer_size, in_features, out_features, group_size, stream ),
            "cuda_rmsnorm_fp4a16_gemm" );
    }

    // ---- RMSNorm -> gate|up Linear -> GLU as one call: ln_2 -> fc_gate_up -> geglu / swiglu (Gemma.Block.ixx:209-210,347-349) ----
    inline void cuda_rmsnorm_w8a16_gemm_glu(
        __nv_bfloat16* output, __nv_bfloat16* gate_up_scratch, __nv_bfloat16* normed_scratch, const __nv_bfloat16* activations,
        const __nv_bfloat16* norm_weight, const __nv_bfloat16* norm_bias, float epsilon, float weight_offset,
        const __nv_fp8_e4m3* weight, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int hidden, int glu_kind, cudaStream_t stream )
    {
        milab200_detail::check( milab200_rmsnorm_w8a16_gemm_glu( output, gate_up_scratch, normed_scratch, activations, norm_weight,
            norm_bias, epsilon, weight_offset, weight, scales, bias, outer_size, in_features, hidden, glu_kind, stream ),
            "cuda_rmsnorm_w8a16_gemm_glu" );
    }
    inline void cuda_rmsnorm_fp4a16_gemm_glu(
        __nv_bfloat16* output, __nv_bfloat16* gate_up_scratch, __nv_bfloat16* normed_scratch, const __nv_bfloat16* activations,
        const __nv_bfloat16* norm_weight, const __nv_bfloat16* norm_bias, float epsilon, float weight_offset,
        const uint8_t* weights_packed, const float* scales, const __nv_bfloat16* bias,
        int outer_size, int in_features, int hidden, int group_size, int glu_kind, cudaStream_t stream )
    {
        milab200_detail::check( milab200_rmsnorm_fp4a16_gemm_glu( output, gate_up_scratch, normed_scratch, activations, norm_weight,
            norm_bias, epsilon, weight_offset, weights_packed, scales, bias, outer_size, in_features, hidden, group_size, glu_kind,
            stream ), "cuda_rmsnorm_fp4a16_gemm_glu" );
    }
}

// ---- the launchers of the neighbouring ops this library also replaces bit for bit -----------------------------
namespace Mila::Dnn::Compute::Cuda::TokenEmbedding
{
    // Embeddings/Kernels/TokenEmbedding.cuh:46-52
    inline void cuda_token_embedding_forward_bf16_qfp8(
        __nv_bfloat16* Y, const int* X, const void* wte_fp8, const float* scales, int B, int T, int C, cudaStream_t stream )
    {
        Mila::Dnn::Compute::Cuda::Linear::milab200_detail::check(
            milab200_token_embedding_forward_bf16_qfp8( Y, X, wte_fp8, scales, B, T, C, stream ), "cuda_token_embedding_forward_bf16_qfp8" );
    }
    inline void cuda_token_embedding_decode_bf16_qfp8(
        __nv_bfloat16* Y, const int* X, const void* wte_fp8, const float* scales, int B, int C, cudaStream_t stream )
    {
        Mila::Dnn::Compute::Cuda::Linear::milab200_detail::check(
            milab200_token_embedding_decode_bf16_qfp8( Y, X, wte_fp8, scales, B, C, stream ), "cuda_token_embedding_decode_bf16_qfp8" );
    }
}
namespace Mila::Dnn::Compute::Cuda::Geglu
{
    // Activations/Geglu/Kernels/Geglu.cuh
    inline void cuda_geglu_forward_bf16( __nv_bfloat16* Y, const __nv_bfloat16* X, int N, int half_width, cudaStream_t stream )
    {
        Mila::Dnn::Compute::Cuda::Linear::milab200_detail::check(
            milab200_geglu_forward_bf16( Y, X, N, half_width, stream ), "cuda_geglu_forward_bf16" );
    }
}
namespace Mila::Dnn::Compute::Cuda::Swiglu
{
    // Activations/Swiglu/Kernels/Swiglu.cuh:30-34
    inline void cuda_swiglu_forward_bf16( __nv_bfloat16* Y, const __nv_bfloat16* X, int N, int half_width, cudaStream_t stream )
    {
        Mila::Dnn::Compute::Cuda::Linear::milab200_detail::check(
            milab200_swiglu_forward_bf16( Y, X, N, half_width, stream ), "cuda_swiglu_forward_bf16" );
    }
}
namespace Mila::Dnn::Compute::Cuda::RmsNorm
{
    // Normalizations/RmsNorm/Kernels/RmsNorm.cuh:125
    inline void cuda_rmsnorm_forward_bf16(
        __nv_bfloat16* Y, __nv_bfloat16* rstd, const __nv_bfloat16* X, const __nv_bfloat16* weight, const __nv_bfloat16* bias,
        int outer_size, int inner_size, int norm_dim, float epsilon, float weight_offset, cudaStream_t stream )
    {
        Mila::Dnn::Compute::Cuda::Linear::milab200_detail::check(
            milab200_rmsnorm_forward_bf16( Y, rstd, X, weight, bias, outer_size, inner_size, norm_dim, epsilon, weight_offset, stream ),
            "cuda_rmsnorm_forward_bf16" );
    }
}
