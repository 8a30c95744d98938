// embed.cu — FP8 tied-table token-embedding gather-dequant (the embedding that shares the
// lm_head's PerChannelFp8 table and scales, Gemma.ixx:139-147,:540-556).
//
// Replaces cuda_token_embedding_forward_bf16_qfp8 / cuda_token_embedding_decode_bf16_qfp8
// (Embeddings/Kernels/TokenEmbedding.cuh:46-52, impl TokenEmbedding.Fp8.cu:34-66,:71-101,:107-147):
//   Y[bt, :] = bf16( f32(W8[X[bt], :]) * scales[X[bt]] )            (RN to BF16, one multiply: bit-exact)
// The table itself is produced by the PerChannelFp8 quantizer of this library, called per row chunk exactly as
// CudaTokenEmbeddingOp.Quantize.ixx:63-113 calls the reference's (chunk_rows as out_features).
// HBM-bound and tiny (B*T*C bytes in, 2*B*T*C out): 16 table bytes -> 32 output bytes per thread, int64 offsets
// (a 262144 x 3840 table has ~1e9 elements).
#include "common.cuh"

namespace milab200 {
namespace {

template <int VEC>      // table bytes per thread: 16 when C % 16 == 0, else 8 (the reference requires C % 8 == 0)
__global__ void __launch_bounds__(256)
token_embedding_qfp8_kernel(__nv_bfloat16* __restrict__ Y, const int* __restrict__ X, const uint8_t* __restrict__ W8,
                            const float* __restrict__ scales, long long rows, int C)
{
    const int cv = C / VEC;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cv) return;
    const long long bt = idx / cv;
    const int c = (int)(idx - bt * cv) * VEC;
    const int ix = __ldg(X + bt);
    const float s = __ldg(scales + ix);
    const uint8_t* src = W8 + (long long)ix * C + c;
    uint32_t w[VEC / 4];
    if constexpr (VEC == 16) { const uint4 r = ldg_stream_v4(src); w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w; }
    else                     { const uint2 r = ldg_stream_v2(src); w[0] = r.x; w[1] = r.y; }
    uint32_t o[VEC / 2];
#pragma unroll
    for (int j = 0; j < VEC / 4; ++j) {
        uint32_t lo, hi;
        cvt_e4m3x4_to_f16x2x2(w[j], lo, hi);                    // exact: E4M3 is a subset of F16
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x * s, a.y * s);
        const __nv_bfloat162 p1 = __floats2bfloat162_rn(b.x * s, b.y * s);
        o[2 * j] = *reinterpret_cast<const uint32_t*>(&p0);
        o[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&p1);
    }
    __nv_bfloat16* dst = Y + bt * C + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    if constexpr (VEC == 16) *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
}

}  // namespace

int launch_token_embedding_qfp8(void* Y, const int* X, const void* wte_fp8, const float* scales, long long rows, int C,
                                cudaStream_t stream)
{
    if (!Y || !X || !wte_fp8 || !scales || rows <= 0 || C <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (C % 8 != 0) return MILAB200_E_BAD_SHAPE;
    auto* y = static_cast<__nv_bfloat16*>(Y);
    auto* w = static_cast<const uint8_t*>(wte_fp8);
    if (C % 16 == 0) {
        const long long n = rows * (C / 16);
        token_embedding_qfp8_kernel<16><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(y, X, w, scales, rows, C);
    } else {
        const long long n = rows * (C / 8);
        token_embedding_qfp8_kernel<8><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(y, X, w, scales, rows, C);
    }
    note_launch("token_embedding_qfp8_kernel");
    return (int)cudaGetLastError();
}

}  // namespace milab200

using namespace milab200;

extern "C" {

int milab200_token_embedding_forward_bf16_qfp8(void* Y, const int* X, const void* wte_fp8, const float* scales,
                                               int B, int T, int C, milab200_stream_t stream)
{
    if (B <= 0 || T <= 0) return MILAB200_E_INVALID_ARGUMENT;
    return launch_token_embedding_qfp8(Y, X, wte_fp8, scales, (long long)B * T, C, static_cast<cudaStream_t>(stream));
}

int milab200_token_embedding_decode_bf16_qfp8(void* Y, const int* X, const void* wte_fp8, const float* scales,
                                              int B, int C, milab200_stream_t stream)
{
    return launch_token_embedding_qfp8(Y, X, wte_fp8, scales, B, C, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
