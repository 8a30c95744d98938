// gemm.cu — batched (M > 16) Linear forward for FP8-per-channel and FP4-per-group weights.
//
// Replaces cuda_w8a16_gemm (LIN/Kernels/W8A16Gemm/CudaW8A16Gemm.cu:62-162) and
// cuda_fp4a16_gemm / cuda_fp4a16_gemm_wmma (LIN/Kernels/W4A16Gemm/CudaW4A16Gemm.cu:88-400,
// CudaW4A16Gemm.Wmma.cu:145-333).
//
// Round-1 state: token-blocked streaming — the decode kernel (gemv.cu) is run over blocks of 16
// tokens, so every block re-streams the weights (from L2 when the layer fits).  Numerically this
// is exactly the decode path (FP32-exact weights, FP32 accumulate).  The TMA + tcgen05/TMEM
// prefill kernel that replaces it for M >= 128 is the next kernel on this file's list
// (DESIGN.md "Prefill").
#include "common.cuh"

namespace milab200 {

int launch_gemv_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemv_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);

int launch_gemm_fp8(void* out, const void* act, const void* w, const float* scales, const void* bias,
                    int M, int K, int N, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    auto* o = static_cast<__nv_bfloat16*>(out);
    auto* a = static_cast<const __nv_bfloat16*>(act);
    for (int m0 = 0; m0 < M; m0 += 16) {
        const int mb = (M - m0 < 16) ? (M - m0) : 16;
        const int rc = launch_gemv_fp8(o + (size_t)m0 * N, a + (size_t)m0 * K, w, scales, bias, mb, K, N, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

int launch_gemm_fp4(void* out, const void* act, const void* w, const float* scales, const void* bias,
                    int M, int K, int N, int group_size, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    auto* o = static_cast<__nv_bfloat16*>(out);
    auto* a = static_cast<const __nv_bfloat16*>(act);
    for (int m0 = 0; m0 < M; m0 += 16) {
        const int mb = (M - m0 < 16) ? (M - m0) : 16;
        const int rc = launch_gemv_fp4(o + (size_t)m0 * N, a + (size_t)m0 * K, w, scales, bias, mb, K, N, group_size, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

}  // namespace milab200
