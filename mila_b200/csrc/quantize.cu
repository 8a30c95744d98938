// quantize.cu — load-time weight quantizers (PerChannelFp8<>, PerGroupFp4<g>) for sm_100a.
//
// Bit-exact with Mila's kernels (format spec: SURVEY.md §2.3):
//   FP8: LIN/Kernels/Quantization/CudaFp8WeightQuantization.cu:57-121
//   FP4: LIN/Kernels/Quantization/CudaFp4WeightQuantization.cu:54-70, 83-144
// The arithmetic that decides bits is kept identical (fmaxf absmax, `absmax / C`, `1.0f / scale`
// as two IEEE divisions, one FP32 multiply, cvt.rn.satfinite.e4m3x2.f32 / the strict-less-than
// E2M1 ladder); everything else is redesigned for B200: 128-bit loads, each BF16 element is
// read from HBM once, FP4 is a flat warp-per-256-elements streaming kernel (shuffle max, 32-bit
// packed stores) instead of one 128-thread block with two barriers per group.
#include "common.cuh"

namespace milab200 {

// ------------------------------------------------------------------------------------------
// PerChannelFp8: one CTA per output channel; the row lives in registers between the absmax
// pass and the convert pass (up to kCache * 256 * 8 = 16384 columns; longer rows re-read the
// tail from L2).
// ------------------------------------------------------------------------------------------
namespace {

constexpr int kQ8Threads = 256;
constexpr int kQ8Cache = 8;   // uint4 (8 x bf16) chunks cached per thread

__device__ __forceinline__ float absmax8(const uint4& v, float m)
{
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m = fmaxf(m, fabsf(bf16lo(w[i])));
        m = fmaxf(m, fabsf(bf16hi(w[i])));
    }
    return m;
}

__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d)
{
    // cvt.rn.satfinite.e4m3x2.f32 d, hi, lo  — the same instruction `__nv_fp8_e4m3(float)` emits.
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(lo) : "f"(a), "f"(b));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(hi) : "f"(c), "f"(d));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

__device__ __forceinline__ uint2 quant8(const uint4& v, float inv)
{
    uint2 o;
    o.x = e4m3x4(bf16lo(v.x) * inv, bf16hi(v.x) * inv, bf16lo(v.y) * inv, bf16hi(v.y) * inv);
    o.y = e4m3x4(bf16lo(v.z) * inv, bf16hi(v.z) * inv, bf16lo(v.w) * inv, bf16hi(v.w) * inv);
    return o;
}

__global__ void __launch_bounds__(kQ8Threads)
quantize_fp8_per_channel_kernel(const __nv_bfloat16* __restrict__ src, uint8_t* __restrict__ dst,
                                float* __restrict__ scales, int64_t K)
{
    __shared__ float s_red[kQ8Threads / 32];
    const int64_t row = blockIdx.x;
    const __nv_bfloat16* rsrc = src + row * K;
    uint8_t* rdst = dst + row * K;
    const int tid = threadIdx.x;

    // vector path needs 16-byte aligned rows: K % 8 == 0 (base pointers are cudaMalloc-aligned)
    const bool vec = (K % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
    const int64_t nchunks = vec ? K / 8 : 0;

    uint4 cache[kQ8Cache];
    float m = 0.0f;
    if (vec) {
#pragma unroll
        for (int i = 0; i < kQ8Cache; ++i) {
            const int64_t c = tid + (int64_t)i * kQ8Threads;
            if (c < nchunks) { cache[i] = ldg_stream_v4(rsrc + c * 8); m = absmax8(cache[i], m); }
        }
        for (int64_t c = tid + (int64_t)kQ8Cache * kQ8Threads; c < nchunks; c += kQ8Threads)
            m = absmax8(ldg_cached_v4(rsrc + c * 8), m);
    } else {
        for (int64_t k = tid; k < K; k += kQ8Threads)
            m = fmaxf(m, fabsf(__bfloat162float(rsrc[k])));
    }

    m = warp_max(m);
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    __syncthreads();
    float absmax = s_red[0];
#pragma unroll
    for (int w = 1; w < kQ8Threads / 32; ++w) absmax = fmaxf(absmax, s_red[w]);

    const float scale = (absmax > 0.0f) ? __fdiv_rn(absmax, 448.0f) : 1.0f;
    const float inv = __fdiv_rn(1.0f, scale);
    if (tid == 0) scales[row] = scale;

    if (vec) {
#pragma unroll
        for (int i = 0; i < kQ8Cache; ++i) {
            const int64_t c = tid + (int64_t)i * kQ8Threads;
            if (c < nchunks) *reinterpret_cast<uint2*>(rdst + c * 8) = quant8(cache[i], inv);
        }
        for (int64_t c = tid + (int64_t)kQ8Cache * kQ8Threads; c < nchunks; c += kQ8Threads)
            *reinterpret_cast<uint2*>(rdst + c * 8) = quant8(ldg_cached_v4(rsrc + c * 8), inv);
    } else {
        for (int64_t k = tid; k < K; k += kQ8Threads) {
            uint16_t p;
            const float v = __bfloat162float(rsrc[k]) * inv;
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(p) : "f"(v), "f"(0.0f));
            rdst[k] = (uint8_t)(p & 0xFFu);
        }
    }
}

// ------------------------------------------------------------------------------------------
// PerGroupFp4<g>: flat streaming kernel.  Because K % g == 0 the [N,K] matrix is a flat array
// of N*K/g groups; lane l of a warp owns 8 consecutive elements (one 128-bit load), g/8 lanes
// share a group.  absmax by xor-shuffle inside the lane group; 8 nibbles -> one 32-bit store.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t e2m1_encode(float x)
{
    // strict-less-than ladder, written as a count of failed `<` tests so NaN -> 7 exactly like
    // the reference's if/else chain; sign test is `x < 0` (so -0.0f -> 0, NaN -> +).
    const float a = fabsf(x);
    uint32_t mag = 0;
    mag += !(a < 0.25f);
    mag += !(a < 0.75f);
    mag += !(a < 1.25f);
    mag += !(a < 1.75f);
    mag += !(a < 2.5f);
    mag += !(a < 3.5f);
    mag += !(a < 5.0f);
    return mag | ((x < 0.0f) ? 8u : 0u);
}

template <int G>
__global__ void __launch_bounds__(256)
quantize_fp4_per_group_kernel(const __nv_bfloat16* __restrict__ src, uint32_t* __restrict__ dst_words,
                              float* __restrict__ scales, int64_t total_chunks /* N*K/8 */)
{
    constexpr int kLanesPerGroup = G / 8;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);

    for (int64_t base = warp * 32; base < total_chunks; base += nwarps * 32) {
        const int64_t c = base + lane;
        const bool live = c < total_chunks;        // whole lane groups are live or dead together
        uint4 v = make_uint4(0, 0, 0, 0);
        if (live) v = ldg_stream_v4(src + c * 8);

        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { f[2 * i] = bf16lo(w[i]); f[2 * i + 1] = bf16hi(w[i]); }

        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(f[i]));
#pragma unroll
        for (int s = kLanesPerGroup / 2; s > 0; s >>= 1)
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));

        const float scale = (m > 0.0f) ? __fdiv_rn(m, 6.0f) : 1.0f;
        const float inv = __fdiv_rn(1.0f, scale);

        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) word |= e2m1_encode(f[i] * inv) << (4 * i);

        if (live) {
            dst_words[c] = word;
            if ((lane % kLanesPerGroup) == 0) scales[c / kLanesPerGroup] = scale;
        }
    }
}

int grid_for_stream_kernel(int64_t warps_needed)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ctas_needed = (warps_needed + 7) / 8;
    const int64_t cap = (int64_t)sms * 8;          // 8 resident 256-thread CTAs per SM
    return (int)(ctas_needed < cap ? (ctas_needed > 0 ? ctas_needed : 1) : cap);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------

int launch_quantize_fp8_per_channel(const void* src_dev, void* dst, float* scales,
                                    int64_t N, int64_t K, cudaStream_t stream)
{
    if (!src_dev || !dst || !scales || N <= 0 || K <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (N > 0x7FFFFFFFLL) return MILAB200_E_INVALID_ARGUMENT;
    quantize_fp8_per_channel_kernel<<<(unsigned)N, kQ8Threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(src_dev), static_cast<uint8_t*>(dst), scales, K);
    note_launch("quantize_fp8_per_channel_kernel");
    return (int)cudaGetLastError();
}

int launch_quantize_fp4_per_group(const void* src_dev, void* dst_packed, float* scales,
                                  int64_t N, int64_t K, int group_size, cudaStream_t stream)
{
    if (!src_dev || !dst_packed || !scales || N <= 0 || K <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    if (K % group_size != 0) return MILAB200_E_BAD_SHAPE;
    const int64_t chunks = N * K / 8;
    const int grid = grid_for_stream_kernel((chunks + 31) / 32);
    if (group_size == 128)
        quantize_fp4_per_group_kernel<128><<<grid, 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(src_dev), static_cast<uint32_t*>(dst_packed), scales, chunks);
    else
        quantize_fp4_per_group_kernel<64><<<grid, 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(src_dev), static_cast<uint32_t*>(dst_packed), scales, chunks);
    note_launch(group_size == 128 ? "quantize_fp4_per_group_kernel<128>" : "quantize_fp4_per_group_kernel<64>");
    return (int)cudaGetLastError();
}

}  // namespace milab200
