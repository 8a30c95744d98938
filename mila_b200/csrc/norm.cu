// norm.cu — RMSNorm -> Linear as one call (second half) and the stand-alone BF16 RMSNorm.
//
// milab200_rmsnorm_forward_bf16 replaces cuda_rmsnorm_forward_bf16 (Normalizations/RmsNorm/Kernels/RmsNorm.cuh:125,
// kernel RmsNorm.Bf16.cu:19-73) bit for bit.  milab200_rmsnorm_{w8a16,fp4a16}_gemm run Linear(RMSNorm(x)): on the
// tcgen05 routes the normalisation happens in the activation converters / pre-pass (norm.cuh) and the normalised
// tensor never exists in memory; shapes those routes do not take run the stand-alone kernel into the caller's scratch
// and then the ordinary Linear — the same bits either way.
#include <atomic>

#include "gemv_common.cuh"
#include "glu.cuh"
#include "norm.cuh"

namespace milab200 {
using namespace gemv;

int try_decode_tc_norm(int fmt, __nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                       int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu, const NormArgs* norm);
int try_decode_mx4_norm(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                        int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu, const NormArgs* norm);      // decode_mx4.cu
int launch_linear_glu(int fmt, void* out, void* gate_up_scratch, const void* act, const void* w, const float* scales,
                      const void* bias, int M, int K, int H, int kind, cudaStream_t stream);                        // glu.cu
int try_prefill_tc_glu(int fmt, __nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                       int M, int K, int H, int glu_kind, cudaStream_t, int* status, const NormArgs* norm);
int try_prefill_tc_norm(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                        const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status, const NormArgs* norm);
int launch_gemv_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemv_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int launch_gemm_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemm_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);

// "rmsnorm_fast_reduction" (milab200_set_option): the fused RMSNorm -> Linear routes compute the reciprocal RMS in tree order
// (norm.cuh NormArgs::fast).  Off by default: the default is bit-identical to the reference's kernel sequence.
std::atomic<int> g_norm_fast{ 0 };
void norm_set_fast(int on) { g_norm_fast.store(on != 0 ? 1 : 0); }

namespace {

// One warp per normalisation slice; slices are strided by inner_size along the normalised axis.
__global__ void __launch_bounds__(512)
rmsnorm_forward_bf16_kernel(__nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ rstd_out,
                            const __nv_bfloat16* __restrict__ inp, const __nv_bfloat16* __restrict__ weight,
                            const __nv_bfloat16* __restrict__ bias, int num_slices, int norm_dim, int inner_size,
                            float eps, float weight_offset)
{
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= num_slices) return;
    const size_t base = (size_t)(idx / inner_size) * norm_dim * inner_size + (idx % inner_size);
    const __nv_bfloat16* x = inp + base;
    __nv_bfloat16* o = out + base;
    float rstd;
    if (inner_size == 1) rstd = rms_rstd_warp(x, norm_dim, eps, lane);
    else {
        float m2 = 0.0f;
        for (int i = lane; i < norm_dim; i += 32) {
            const float v = __bfloat162float(x[(size_t)i * inner_size]);
            m2 = fmaf(v, v, m2);
        }
#pragma unroll
        for (int offset = 16; offset > 0; offset >>= 1) m2 += __shfl_down_sync(0xffffffffu, m2, offset);
        m2 = __shfl_sync(0xffffffffu, m2, 0);
        rstd = rsqrtf(__fdiv_rn(m2, (float)norm_dim) + eps);
    }
    if (lane == 0 && rstd_out) rstd_out[idx] = __float2bfloat16(rstd);
    for (int i = lane; i < norm_dim; i += 32) {
        const size_t off = (size_t)i * inner_size;
        const float w = weight ? __fadd_rn(__bfloat162float(weight[i]), weight_offset) : 1.0f;
        const float b = bias ? __bfloat162float(bias[i]) : 0.0f;
        o[off] = __float2bfloat16(rms_apply1(__bfloat162float(x[off]), rstd, w, b));
    }
}

int rmsnorm_launch(void* Y, void* rstd, const void* X, const void* weight, const void* bias, int outer, int inner, int norm_dim,
                   float eps, float weight_offset, cudaStream_t stream)
{
    if (!Y || !X || outer <= 0 || inner <= 0 || norm_dim <= 0) return MILAB200_E_INVALID_ARGUMENT;
    const int slices = outer * inner, wpb = 512 / 32;
    rmsnorm_forward_bf16_kernel<<<(slices + wpb - 1) / wpb, 512, 0, stream>>>(
        static_cast<__nv_bfloat16*>(Y), static_cast<__nv_bfloat16*>(rstd), static_cast<const __nv_bfloat16*>(X),
        static_cast<const __nv_bfloat16*>(weight), static_cast<const __nv_bfloat16*>(bias), slices, norm_dim, inner, eps, weight_offset);
    note_launch("rmsnorm_forward_bf16_kernel");
    return (int)cudaGetLastError();
}

int rmsnorm_linear(int fmt, void* out, void* normed_scratch, const void* act, const void* norm_weight, const void* norm_bias,
                   float eps, float weight_offset, const void* w, const float* scales, const void* bias,
                   int M, int K, int N, int group_size, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    NormArgs na;
    na.weight = static_cast<const __nv_bfloat16*>(norm_weight); na.bias = static_cast<const __nv_bfloat16*>(norm_bias);
    na.eps = eps; na.weight_offset = weight_offset; na.on = 1;
    na.fast = (g_norm_fast.load(std::memory_order_relaxed)) ? 1 : 0;
    auto* o = static_cast<__nv_bfloat16*>(out);
    auto* a = static_cast<const __nv_bfloat16*>(act);
    auto* W = static_cast<const uint8_t*>(w);
    auto* B = static_cast<const __nv_bfloat16*>(bias);
    int status = 0;
    if (K % 8 == 0 && (fmt == kFp8 || fmt == kFp4G128)) {
        if (M <= kMaxTok) {
            if (fmt == kFp4G128 && try_decode_mx4_norm(o, a, W, scales, B, M, K, N, stream, &status, nullptr, 0, &na) == 0) return status;
            if (try_decode_tc_norm(fmt, o, a, W, scales, B, M, K, N, stream, &status, nullptr, 0, &na) == 0) return status;
        } else if (M > 32 && try_prefill_tc_norm(fmt, o, a, W, scales, B, M, K, N, stream, &status, &na) == 0) return status;
        // (16 < M <= 32 is routed per layer shape between two kernels, gemm.cu: those calls take the two-kernel sequence)
    }
    // not a fused route: the stand-alone kernel into the caller's scratch, then the ordinary Linear
    if (!normed_scratch) return MILAB200_E_INVALID_ARGUMENT;
    const int rc = rmsnorm_launch(normed_scratch, nullptr, act, norm_weight, norm_bias, M, 1, K, eps, weight_offset, stream);
    if (rc != 0) return rc;
    if (fmt == kFp8)
        return (M <= kMaxTok) ? launch_gemv_fp8(out, normed_scratch, w, scales, bias, M, K, N, stream)
                              : launch_gemm_fp8(out, normed_scratch, w, scales, bias, M, K, N, stream);
    return (M <= kMaxTok) ? launch_gemv_fp4(out, normed_scratch, w, scales, bias, M, K, N, group_size, stream)
                          : launch_gemm_fp4(out, normed_scratch, w, scales, bias, M, K, N, group_size, stream);
}

// RMSNorm -> gate|up Linear -> GLU: the whole front half of Mila's MLP block (ln_2 -> fc_gate_up -> geglu / swiglu,
// Gemma.Block.ixx:209-210,347-349; Llama.Block.ixx:883) as ONE launch on the decode routes — the norm in the activation
// converters, the gated activation in the epilogue.  Bit for bit the three-kernel sequence.
int rmsnorm_linear_glu(int fmt, void* out, void* gate_up_scratch, void* normed_scratch, const void* act, const void* norm_weight,
                       const void* norm_bias, float eps, float weight_offset, const void* w, const float* scales, const void* bias,
                       int M, int K, int H, int kind, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || H <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (kind != kGluGegluTanh && kind != kGluSwiglu) return MILAB200_E_INVALID_ARGUMENT;
    NormArgs na;
    na.weight = static_cast<const __nv_bfloat16*>(norm_weight); na.bias = static_cast<const __nv_bfloat16*>(norm_bias);
    na.eps = eps; na.weight_offset = weight_offset; na.on = 1;
    na.fast = (g_norm_fast.load(std::memory_order_relaxed)) ? 1 : 0;
    if (M <= kMaxTok && K % 128 == 0 && (fmt == kFp8 || fmt == kFp4G128)) {
        int status = 0;
        auto* y = static_cast<__nv_bfloat16*>(out);
        auto* x = static_cast<const __nv_bfloat16*>(act);
        auto* W = static_cast<const uint8_t*>(w);
        auto* B = static_cast<const __nv_bfloat16*>(bias);
        if (fmt == kFp4G128 && try_decode_mx4_norm(y, x, W, scales, B, M, K, 2 * H, stream, &status, nullptr, kind, &na) == 0) return status;
        if (try_decode_tc_norm(fmt, y, x, W, scales, B, M, K, 2 * H, stream, &status, nullptr, kind, &na) == 0) return status;
    }
    if (M > 32 && K % 128 == 0 && (fmt == kFp8 || fmt == kFp4G128)) {
        // batched: the norm inside the activation pre-pass, the gated activation in the GEMM epilogue — one pre-pass + one GEMM
        int status = 0;
        if (try_prefill_tc_glu(fmt, static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(act), static_cast<const uint8_t*>(w),
                               scales, static_cast<const __nv_bfloat16*>(bias), M, K, H, kind, stream, &status, &na) == 0) return status;
    }
    // not a fused route: the stand-alone norm into the caller's scratch, then the gate|up Linear + GLU
    if (!normed_scratch) return MILAB200_E_INVALID_ARGUMENT;
    const int rc = rmsnorm_launch(normed_scratch, nullptr, act, norm_weight, norm_bias, M, 1, K, eps, weight_offset, stream);
    if (rc != 0) return rc;
    return launch_linear_glu(fmt, out, gate_up_scratch, normed_scratch, w, scales, bias, M, K, H, kind, stream);
}

}  // namespace
}  // namespace milab200

using namespace milab200;

extern "C" {

int milab200_rmsnorm_forward_bf16(void* Y, void* rstd, const void* X, const void* weight, const void* bias,
                                  int outer_size, int inner_size, int norm_dim, float epsilon, float weight_offset,
                                  milab200_stream_t stream)
{
    return rmsnorm_launch(Y, rstd, X, weight, bias, outer_size, inner_size, norm_dim, epsilon, weight_offset, static_cast<cudaStream_t>(stream));
}

int milab200_rmsnorm_w8a16_gemm(void* out, void* normed_scratch, const void* act, const void* norm_weight, const void* norm_bias,
                                float epsilon, float weight_offset, const void* w, const float* scales, const void* bias,
                                int M, int K, int N, milab200_stream_t stream)
{
    return rmsnorm_linear(kFp8, out, normed_scratch, act, norm_weight, norm_bias, epsilon, weight_offset, w, scales, bias, M, K, N, 0,
                          static_cast<cudaStream_t>(stream));
}

int milab200_rmsnorm_fp4a16_gemm(void* out, void* normed_scratch, const void* act, const void* norm_weight, const void* norm_bias,
                                 float epsilon, float weight_offset, const void* w, const float* scales, const void* bias,
                                 int M, int K, int N, int group_size, milab200_stream_t stream)
{
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    return rmsnorm_linear(group_size == 128 ? kFp4G128 : kFp4G64, out, normed_scratch, act, norm_weight, norm_bias, epsilon, weight_offset,
                          w, scales, bias, M, K, N, group_size, static_cast<cudaStream_t>(stream));
}


int milab200_rmsnorm_w8a16_gemm_glu(void* out, void* gate_up_scratch, void* normed_scratch, const void* act, const void* norm_weight,
                                    const void* norm_bias, float epsilon, float weight_offset, const void* w, const float* scales,
                                    const void* bias, int M, int K, int H, int glu_kind, milab200_stream_t stream)
{
    return rmsnorm_linear_glu(kFp8, out, gate_up_scratch, normed_scratch, act, norm_weight, norm_bias, epsilon, weight_offset, w, scales,
                              bias, M, K, H, glu_kind, static_cast<cudaStream_t>(stream));
}

int milab200_rmsnorm_fp4a16_gemm_glu(void* out, void* gate_up_scratch, void* normed_scratch, const void* act, const void* norm_weight,
                                     const void* norm_bias, float epsilon, float weight_offset, const void* w, const float* scales,
                                     const void* bias, int M, int K, int H, int group_size, int glu_kind, milab200_stream_t stream)
{
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    return rmsnorm_linear_glu(group_size == 128 ? kFp4G128 : kFp4G64, out, gate_up_scratch, normed_scratch, act, norm_weight, norm_bias,
                              epsilon, weight_offset, w, scales, bias, M, K, H, glu_kind, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
