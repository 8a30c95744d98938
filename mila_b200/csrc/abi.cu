// abi.cu — extern "C" surface of libmila_b200_linear.so (declared in include/mila_b200_linear.h).
// Argument checking, the H2D staging copy of the quantizers, M-regime routing.  No entry point
// synchronises, allocates, or falls back to the CPU.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace milab200 {

// kernels / launchers defined in the other translation units
int launch_quantize_fp8_per_channel(const void*, void*, float*, int64_t, int64_t, cudaStream_t);
int launch_quantize_fp4_per_group(const void*, void*, float*, int64_t, int64_t, int, cudaStream_t);
int launch_gemv_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemv_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int launch_gemv_generic(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int launch_gemm_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemm_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int launch_fp8_dequantize_to_bf16(void*, const void*, const float*, int, int, cudaStream_t);
int launch_fp4_dequantize_to_bf16(void*, const void*, const float*, int, int, int, cudaStream_t);
int launch_compute_fp8_weight_scale(float*, const float*, int64_t, cudaStream_t);
int launch_fp4_dequantize_to_fp8(void*, const void*, const float*, const float*, int, int, int, cudaStream_t);
int launch_quantize_bf16_to_fp8_per_token(void*, float*, const void*, int, int, cudaStream_t);
int launch_fp8_apply_per_token_scales(void*, const float*, const void*, int, int, cudaStream_t);
int launch_add_bias_bf16(void*, const void*, int, int, cudaStream_t);
int launch_add_bias_f32(float*, const float*, int, int, cudaStream_t);
int tc_prepare_device();
void tc_note_weights_written();
void tc_set_enabled(bool);
void tc_set_prof(long long*);
void tc_set_streamk(int);
void tc_set_presplit(int);
void prefill_tc_set_enabled(bool);
void mx4_set_max_m(int);
void mx4_set_pair(int);
void mx4_set_coop(int);
void prefill_tc_set_cta_group(int);
void prefill_tc_set_fp4_sum(int);
void prefill_tc_set_glu(int);
void norm_set_fast(int);
void prefill_tc_set_planes(int);
void gemv_set_force_generic(bool);
int prefill_tc_reserve(int, int);

static std::atomic<uint64_t> g_launches{0};
static thread_local const char* g_last_kernel = "";

void note_launch(const char* kernel_name, uint64_t n)
{
    g_launches.fetch_add(n, std::memory_order_relaxed);
    g_last_kernel = kernel_name;
}

static inline cudaStream_t S(milab200_stream_t s) { return static_cast<cudaStream_t>(s); }

constexpr int kDecodeMaxM = 16;

}  // namespace milab200

using namespace milab200;

extern "C" {

int milab200_abi_version(void) { return 1; }

int milab200_init(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return MILAB200_E_NO_DEVICE; }
    return tc_prepare_device();
}

int milab200_reserve_prefill(int max_tokens, int max_in_features)
{
    return prefill_tc_reserve(max_tokens, max_in_features);
}

const char* milab200_error_string(int code)
{
    switch (code) {
        case MILAB200_OK:                  return "success";
        case MILAB200_E_INVALID_ARGUMENT:  return "milab200: invalid argument (null pointer or non-positive size)";
        case MILAB200_E_UNSUPPORTED_GROUP: return "milab200: unsupported group_size (must be 64 or 128)";
        case MILAB200_E_BAD_SHAPE:         return "milab200: in_features must be divisible by group_size and by 8";
        case MILAB200_E_NO_DEVICE:         return "milab200: no usable CUDA device (sm_100 required)";
        case MILAB200_E_NO_NCCL:           return "milab200: NCCL is not loaded in this process or an NCCL call failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "milab200: unknown error";
}

uint64_t milab200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void milab200_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
const char* milab200_last_kernel(void) { return g_last_kernel; }

// ---- quantizers ---------------------------------------------------------------------------

int milab200_quantize_fp8_per_channel(const void* src_bf16_host, void* dst_fp8, float* dst_scales,
                                      int64_t N, int64_t K, void* dev_staging, milab200_stream_t stream)
{
    if (!src_bf16_host || !dev_staging || !dst_fp8 || !dst_scales || N <= 0 || K <= 0)
        return MILAB200_E_INVALID_ARGUMENT;
    const size_t bytes = (size_t)N * (size_t)K * sizeof(__nv_bfloat16);
    MILAB200_RETURN_IF_CUDA(cudaMemcpyAsync(dev_staging, src_bf16_host, bytes, cudaMemcpyHostToDevice, S(stream)));
    tc_note_weights_written();
    return launch_quantize_fp8_per_channel(dev_staging, dst_fp8, dst_scales, N, K, S(stream));
}

int milab200_quantize_fp4_per_group(const void* src_bf16_host, void* dst_packed, float* dst_scales,
                                    int64_t N, int64_t K, int group_size, void* dev_staging,
                                    milab200_stream_t stream)
{
    if (!src_bf16_host || !dev_staging || !dst_packed || !dst_scales || N <= 0 || K <= 0)
        return MILAB200_E_INVALID_ARGUMENT;
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    if (K % group_size != 0) return MILAB200_E_BAD_SHAPE;
    const size_t bytes = (size_t)N * (size_t)K * sizeof(__nv_bfloat16);
    MILAB200_RETURN_IF_CUDA(cudaMemcpyAsync(dev_staging, src_bf16_host, bytes, cudaMemcpyHostToDevice, S(stream)));
    tc_note_weights_written();
    return launch_quantize_fp4_per_group(dev_staging, dst_packed, dst_scales, N, K, group_size, S(stream));
}

int milab200_quantize_fp8_per_channel_device(const void* src, void* dst, float* scales,
                                             int64_t N, int64_t K, milab200_stream_t stream)
{
    tc_note_weights_written();
    return launch_quantize_fp8_per_channel(src, dst, scales, N, K, S(stream));
}

int milab200_quantize_fp4_per_group_device(const void* src, void* dst, float* scales,
                                           int64_t N, int64_t K, int group_size, milab200_stream_t stream)
{
    tc_note_weights_written();
    return launch_quantize_fp4_per_group(src, dst, scales, N, K, group_size, S(stream));
}

// ---- decode ---------------------------------------------------------------------------------

int milab200_matvec_decode_bf16_qfp8(void* y, const void* x, const void* w, const float* scales,
                                     const void* bias, int C, int OC, milab200_stream_t stream)
{
    return launch_gemv_fp8(y, x, w, scales, bias, 1, C, OC, S(stream));
}

int milab200_matvec_decode_bf16_qfp4(void* y, const void* x, const void* w, const float* scales,
                                     const void* bias, int C, int OC, int group_size, milab200_stream_t stream)
{
    return launch_gemv_fp4(y, x, w, scales, bias, 1, C, OC, group_size, S(stream));
}

// ---- batched -------------------------------------------------------------------------------

int milab200_w8a16_gemm(void* out, const void* act, const void* w, const float* scales, const void* bias,
                        int M, int K, int N, milab200_stream_t stream)
{
    if (M <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (M <= kDecodeMaxM) return launch_gemv_fp8(out, act, w, scales, bias, M, K, N, S(stream));
    return launch_gemm_fp8(out, act, w, scales, bias, M, K, N, S(stream));
}

int milab200_fp4a16_gemm(void* out, const void* act, const void* w, const float* scales, const void* bias,
                         int M, int K, int N, int group_size, milab200_stream_t stream)
{
    if (M <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (M <= kDecodeMaxM) return launch_gemv_fp4(out, act, w, scales, bias, M, K, N, group_size, S(stream));
    return launch_gemm_fp4(out, act, w, scales, bias, M, K, N, group_size, S(stream));
}

int milab200_fp4a16_gemm_wmma(void* out, const void* act, const void* w, const float* scales, const void* bias,
                              int M, int K, int N, int group_size, milab200_stream_t stream)
{
    return milab200_fp4a16_gemm(out, act, w, scales, bias, M, K, N, group_size, stream);
}

// ---- staging / W4A8 helpers ------------------------------------------------------------------

int milab200_fp8_dequantize_to_bf16(void* out, const void* w8, const float* scales, int N, int K, milab200_stream_t st)
{ return launch_fp8_dequantize_to_bf16(out, w8, scales, N, K, S(st)); }

int milab200_fp4_dequantize_to_bf16(void* out, const void* packed, const float* scales, int N, int K, int g, milab200_stream_t st)
{ return launch_fp4_dequantize_to_bf16(out, packed, scales, N, K, g, S(st)); }

int milab200_compute_fp8_weight_scale(float* out, const float* group_scales, int64_t n, milab200_stream_t st)
{ return launch_compute_fp8_weight_scale(out, group_scales, n, S(st)); }

int milab200_fp4_dequantize_to_fp8(void* out, const void* packed, const float* scales, const float* sB,
                                   int N, int K, int g, milab200_stream_t st)
{ return launch_fp4_dequantize_to_fp8(out, packed, scales, sB, N, K, g, S(st)); }

int milab200_quantize_bf16_to_fp8_per_token(void* x8, float* sA, const void* x, int M, int K, milab200_stream_t st)
{ return launch_quantize_bf16_to_fp8_per_token(x8, sA, x, M, K, S(st)); }

int milab200_fp8_apply_per_token_scales(void* y, const float* sA, const void* bias, int M, int N, milab200_stream_t st)
{ return launch_fp8_apply_per_token_scales(y, sA, bias, M, N, S(st)); }

int milab200_add_bias_bf16(void* y, const void* bias, int M, int N, milab200_stream_t st)
{ return launch_add_bias_bf16(y, bias, M, N, S(st)); }

int milab200_add_bias_f32(float* y, const float* bias, int M, int N, milab200_stream_t st)
{ return launch_add_bias_f32(y, bias, M, N, S(st)); }

// ---- runtime options (the programmatic form of the MILAB200_* environment switches) ----------
// Route selection only: every route computes the same function and is parity-tested; the defaults are the measured
// best.  Returns MILAB200_E_INVALID_ARGUMENT for an unknown name.
int milab200_set_option(const char* name, int value)
{
    if (!name) return MILAB200_E_INVALID_ARGUMENT;
    if (!std::strcmp(name, "decode_tc"))         { tc_set_enabled(value != 0); return 0; }        // 0: mma.sync decode kernels only
    if (!std::strcmp(name, "decode_mx4_max_m"))  { mx4_set_max_m(value); return 0; }              // largest M of the packed-nibble kernel (default 2)
    if (!std::strcmp(name, "decode_streamk"))    { tc_set_streamk(value); return 0; }             // -1 auto, 0 off, 1 on
    if (!std::strcmp(name, "decode_presplit"))   { tc_set_presplit(value); return 0; }            // M > 8: 1 activation pre-pass, 0 converter warps
    if (!std::strcmp(name, "decode_generic"))    { gemv_set_force_generic(value != 0); return 0; } // 1: one-warp-per-row kernel for every decode call
    if (!std::strcmp(name, "prefill_tc"))        { prefill_tc_set_enabled(value != 0); return 0; } // 0: token-blocked decode kernels for M > 16
    if (!std::strcmp(name, "prefill_cta_group")) { prefill_tc_set_cta_group(value); return 0; }   // 2 CTA pairs (default), 1 single-CTA tiles
    if (!std::strcmp(name, "prefill_act_planes")) { prefill_tc_set_planes(value); return 0; }     // 2 exact split (default), 1 per-token E4M3 (lossy, W4A8-style)
    if (!std::strcmp(name, "prefill_fp4_sum"))   { prefill_tc_set_fp4_sum(value); return 0; }     // FP4 batched: 0 hi|lo columns, 1 summed planes where tiles fill the GPU (default), 2 always
    if (!std::strcmp(name, "prefill_glu"))       { prefill_tc_set_glu(value); return 0; }         // gate|up + GLU for outer_size > 32: 1 activation in the GEMM epilogue (default), 0 Linear + activation kernel
    if (!std::strcmp(name, "rmsnorm_fast_reduction")) { norm_set_fast(value); return 0; }           // fused RMSNorm -> Linear: 1 = tree-order sum of squares (not bit-identical to the reference order), 0 default
    if (!std::strcmp(name, "weights_written"))   { tc_note_weights_written(); return 0; }         // see milab200_note_weights_written
#ifdef MILAB200_DIAG
    if (!std::strcmp(name, "mx8_pair"))          { mx4_set_pair(value); return 0; }
    if (!std::strcmp(name, "mx8_coop"))          { mx4_set_coop(value); return 0; }
#endif
    return MILAB200_E_INVALID_ARGUMENT;
}

// A caller that wrote weight storage with kernels of its own (a device-side copy, a tied-table install) says so:
// the next decode launch on this device then takes plain stream order instead of prefetching weights ahead of
// the previous kernel's completion (programmatic dependent launch).  The library's own quantizers do this themselves.
int milab200_note_weights_written(void) { tc_note_weights_written(); return 0; }

#ifdef MILAB200_DIAG
// bring-up hook (diag build only): device buffer that CTA 0 of the decode kernel fills with role timestamps
void milab200_diag_set_tc_prof(void* buf) { tc_set_prof(static_cast<long long*>(buf)); }
#endif

}  // extern "C"
