// glu.cu — gate|up Linear with the gated activation fused into its epilogue, plus the
// stand-alone GeGLU / SwiGLU forward kernels the unfused fallback (and M > 16) uses.
//
// Reference sequence (Gemma.Block.ixx:347-349, Llama.Block.ixx:883): fc_gate_up_->forward(x) writes
// gate_up [M, 2H] BF16, then cuda_geglu_forward_bf16 (Activations/Geglu/Kernels/Geglu.cu:42-96) or
// cuda_swiglu_forward_bf16 (Activations/Swiglu/Kernels/Swiglu.Bf16.cu:135-230) writes act [M, H].
// At decode the Linear lasts ~10 us and the activation kernel is a second launch plus an [M, 2H] round trip; the
// fused entries do both in the decode GEMV (decode_tc.cu / decode_mx4.cu: a logical tile streams its 128 gate
// rows, then the 128 up rows H below, and combines them in registers).  Same arithmetic as the two-kernel
// sequence: the projections are rounded to BF16 before the activation, the activation is evaluated in FP32 with
// the reference's own expressions (glu.cuh).
#include "gemv_common.cuh"
#include "glu.cuh"
#include "norm.cuh"

namespace milab200 {
using namespace gemv;

int try_decode_tc(int fmt, __nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                  int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu);
int try_decode_mx4(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                   int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu);
int launch_gemv_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemv_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int launch_gemm_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemm_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int try_prefill_tc_glu(int fmt, __nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                       int M, int K, int H, int glu_kind, cudaStream_t, int* status, const NormArgs* norm);

namespace {

// Y[token, col] = glu(X[token, col], X[token, H + col]); 8 BF16 per thread when H % 8 == 0
__global__ void __launch_bounds__(256)
glu_forward_bf16_kernel(__nv_bfloat16* __restrict__ Y, const __nv_bfloat16* __restrict__ X, long long n_out, int H, int kind)
{
    const long long i8 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i8 >= n_out) return;
    if ((H & 7) == 0) {
        const long long token = i8 / H;
        const int col = (int)(i8 - token * H);
        const uint4 g = *reinterpret_cast<const uint4*>(X + token * 2 * H + col);
        const uint4 u = *reinterpret_cast<const uint4*>(X + token * 2 * H + H + col);
        const uint32_t gw[4] = { g.x, g.y, g.z, g.w }, uw[4] = { u.x, u.y, u.z, u.w };
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16 lo = glu_combine(kind, bf16lo(gw[j]), bf16lo(uw[j]));
            const __nv_bfloat16 hi = glu_combine(kind, bf16hi(gw[j]), bf16hi(uw[j]));
            o[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
        }
        *reinterpret_cast<uint4*>(Y + i8) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        for (long long i = i8; i < i8 + 8 && i < n_out; ++i) {
            const long long token = i / H;
            const int col = (int)(i - token * H);
            Y[i] = glu_combine(kind, __bfloat162float(X[token * 2 * H + col]), __bfloat162float(X[token * 2 * H + H + col]));
        }
    }
}

}  // namespace

int launch_glu_forward_bf16(void* Y, const void* X, long long n_out, int H, int kind, cudaStream_t stream)
{
    if (!Y || !X || n_out <= 0 || H <= 0 || n_out % H != 0) return MILAB200_E_INVALID_ARGUMENT;
    if (kind != kGluGegluTanh && kind != kGluSwiglu) return MILAB200_E_INVALID_ARGUMENT;
    const long long threads = (n_out + 7) / 8;
    glu_forward_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        static_cast<__nv_bfloat16*>(Y), static_cast<const __nv_bfloat16*>(X), n_out, H, kind);
    note_launch(kind == kGluGegluTanh ? "glu_forward_bf16_kernel<geglu>" : "glu_forward_bf16_kernel<swiglu>");
    return (int)cudaGetLastError();
}

// fmt: kFp8 / kFp4G128 / kFp4G64.  out [M, H]; W [2H, K]; gate_up_scratch [M, 2H] is the gate|up Linear's own output
// tensor in the reference — written only when the fused kernel does not take the shape.
int launch_linear_glu(int fmt, void* out, void* gate_up_scratch, const void* act, const void* w, const float* scales,
                      const void* bias, int M, int K, int H, int kind, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || H <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (kind != kGluGegluTanh && kind != kGluSwiglu) return MILAB200_E_INVALID_ARGUMENT;
    if (M <= kMaxTok && K % 128 == 0) {
        int status = 0;
        auto* y = static_cast<__nv_bfloat16*>(out);
        auto* x = static_cast<const __nv_bfloat16*>(act);
        auto* W = static_cast<const uint8_t*>(w);
        auto* B = static_cast<const __nv_bfloat16*>(bias);
        if (fmt == kFp4G128 && try_decode_mx4(y, x, W, scales, B, M, K, 2 * H, stream, &status, nullptr, kind) == 0) return status;
        if ((fmt == kFp8 || fmt == kFp4G128) &&
            try_decode_tc(fmt, y, x, W, scales, B, M, K, 2 * H, stream, &status, nullptr, kind) == 0) return status;
    }
    if (M > 32 && K % 128 == 0 && (fmt == kFp8 || fmt == kFp4G128)) {
        // batched: the activation in the epilogue of the TMA + tcgen05 kernel (prefill_tc.cu, CTA pairs: gate rows | up rows)
        int status = 0;
        if (try_prefill_tc_glu(fmt, static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(act), static_cast<const uint8_t*>(w),
                               scales, static_cast<const __nv_bfloat16*>(bias), M, K, H, kind, stream, &status, nullptr) == 0) return status;
    }
    if (!gate_up_scratch) return MILAB200_E_INVALID_ARGUMENT;       // unfused route needs the [M, 2H] tensor
    int rc;
    const int g = (fmt == kFp4G64) ? 64 : 128;
    if (fmt == kFp8) rc = (M <= kMaxTok) ? launch_gemv_fp8(gate_up_scratch, act, w, scales, bias, M, K, 2 * H, stream)
                                         : launch_gemm_fp8(gate_up_scratch, act, w, scales, bias, M, K, 2 * H, stream);
    else             rc = (M <= kMaxTok) ? launch_gemv_fp4(gate_up_scratch, act, w, scales, bias, M, K, 2 * H, g, stream)
                                         : launch_gemm_fp4(gate_up_scratch, act, w, scales, bias, M, K, 2 * H, g, stream);
    if (rc != 0) return rc;
    return launch_glu_forward_bf16(out, gate_up_scratch, (long long)M * H, H, kind, stream);
}

}  // namespace milab200

using namespace milab200;

extern "C" {

int milab200_geglu_forward_bf16(void* Y, const void* X, int N, int half_width, milab200_stream_t stream)
{ return launch_glu_forward_bf16(Y, X, N, half_width, kGluGegluTanh, static_cast<cudaStream_t>(stream)); }

int milab200_swiglu_forward_bf16(void* Y, const void* X, int N, int half_width, milab200_stream_t stream)
{ return launch_glu_forward_bf16(Y, X, N, half_width, kGluSwiglu, static_cast<cudaStream_t>(stream)); }

int milab200_w8a16_gemm_glu(void* out, void* gate_up_scratch, const void* act, const void* w, const float* scales,
                            const void* bias, int M, int K, int H, int glu_kind, milab200_stream_t stream)
{ return launch_linear_glu(gemv::kFp8, out, gate_up_scratch, act, w, scales, bias, M, K, H, glu_kind, static_cast<cudaStream_t>(stream)); }

int milab200_fp4a16_gemm_glu(void* out, void* gate_up_scratch, const void* act, const void* w, const float* scales,
                             const void* bias, int M, int K, int H, int group_size, int glu_kind, milab200_stream_t stream)
{
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    return launch_linear_glu(group_size == 128 ? gemv::kFp4G128 : gemv::kFp4G64, out, gate_up_scratch, act, w, scales, bias,
                             M, K, H, glu_kind, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
