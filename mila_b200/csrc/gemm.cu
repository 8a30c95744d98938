// gemm.cu — batched (M > 16) Linear forward for FP8-per-channel and FP4-per-group weights.
//
// Replaces cuda_w8a16_gemm (LIN/Kernels/W8A16Gemm/CudaW8A16Gemm.cu:62-162) and
// cuda_fp4a16_gemm / cuda_fp4a16_gemm_wmma (LIN/Kernels/W4A16Gemm/CudaW4A16Gemm.cu:88-400,
// CudaW4A16Gemm.Wmma.cu:145-333).
//
// Routing only: M > 32 (M > 16 for wide layers) goes to the TMA + tcgen05/TMEM kernel (prefill_tc.cu) when the shape
// is eligible (K % 128 == 0, FP8 or FP4 g = 128).  Everything else is token-blocked streaming — the
// decode kernel run over blocks of 16 tokens (weights re-streamed from L2 when the layer fits), which
// is numerically exactly the decode path (FP32-exact weights, FP32 accumulate).
#include "gemv_common.cuh"

namespace milab200 {

int launch_gemv_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemv_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int try_prefill_tc(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                   const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status);

#include <cstdlib>
// 16 < M <= 32: the tensor-core kernel wins when the layer has enough 128-row tiles to fill the SMs
// (Llama-8B gate at M = 32: 21 vs 39 us), the token-blocked decode kernel (2 passes, split-K) when it has few
// (Llama-8B down: 44 vs 46 us at M = 32, 43 vs 37 us at M = 17).  Measured on one box, tools/perf_prefill.py.
static bool use_tensor_core_path(int M, int N)
{
    static const int forced = [] { const char* e = std::getenv("MILAB200_BLOCKED_MAX_M"); return (e && *e) ? std::atoi(e) : -1; }();
    if (forced >= 0) return M > forced;
    if (M > 32) return true;
    return M > 16 && (N + 127) / 128 >= 64;
}

int launch_gemm_fp8(void* out, const void* act, const void* w, const float* scales, const void* bias,
                    int M, int K, int N, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    auto* o = static_cast<__nv_bfloat16*>(out);
    auto* a = static_cast<const __nv_bfloat16*>(act);
    if (use_tensor_core_path(M, N)) {
        int status = 0;
        if (try_prefill_tc(gemv::kFp8, o, a, static_cast<const uint8_t*>(w), scales,
                           static_cast<const __nv_bfloat16*>(bias), M, K, N, stream, &status) == 0)
            return status;
    }
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
    if (use_tensor_core_path(M, N) && group_size == 128) {
        int status = 0;
        if (try_prefill_tc(gemv::kFp4G128, o, a, static_cast<const uint8_t*>(w), scales,
                           static_cast<const __nv_bfloat16*>(bias), M, K, N, stream, &status) == 0)
            return status;
    }
    for (int m0 = 0; m0 < M; m0 += 16) {
        const int mb = (M - m0 < 16) ? (M - m0) : 16;
        const int rc = launch_gemv_fp4(o + (size_t)m0 * N, a + (size_t)m0 * K, w, scales, bias, mb, K, N, group_size, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

}  // namespace milab200
