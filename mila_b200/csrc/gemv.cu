// gemv.cu — decode (M = 1..16) Linear forward for FP8-per-channel and FP4-per-group weights.
//
// Replaces cuda_matvec_decode_bf16_qfp8 / _qfp4 (LIN/Kernels/MatVec/CudaMatVecBias.Bf16.cu:198,
// :271, :376) and serves 2 <= M <= 16 of the batched slots.  HBM-bound: the design goal is that
// the SM does (almost) nothing per weight byte except move it.
//
// Why not the reference's structure (one warp per row, FP32 FMA per weight): B200 delivers
// ~23-30 weight bytes per SM per clock; at 2 FP4 weights per byte that is ~50-60 FFMA plus
// ~2.5 decode ops per weight per clock against 128 issue slots — the ALU, not HBM, would bound
// it.  Here instead:
//   * weights go global -> registers with 128-bit streaming loads, 4 steps (1 KB/warp each)
//     in flight per warp;
//   * two packed weights become one f16x2 register with ONE instruction
//     (cvt.rn.f16x2.e2m1x2 / .e4m3x2 -> SASS F2FP.F16.E2M1/E4M3.UNPACK_B) — exact;
//   * the registers are fed as the A fragment of mma.sync.m16n8k16 (16 weight rows x 16 k),
//     the activations (<= 8 tokens per n-tile) are the B fragment, FP32 accumulate in the
//     tensor core.  No per-weight FMA, no LUT, no PRMT.
//   * activations are staged once per CTA in shared memory as FP16 scaled by a per-token
//     power of two (2^-e, e from the token's absmax) so that BF16's range maps into FP16
//     exactly (8-bit significands fit in FP16's 11; every element within 2^-32 of the token
//     max is exact, which is below FP32 accumulation resolution).  They are stored in
//     B-fragment order so the per-step LDS.128 are conflict-free.
//   * FP4 group scales: one k-step == one quantisation group, so the scale multiplies the
//     step's 16x8 partial tile once (4 FFMA per thread per step) — the same factoring the
//     reference's wide kernel uses (Bf16.cu:461-494).
//   * the 8 warps of a CTA split K; partials meet in shared memory once per row tile.
#include <atomic>

#include "gemv_common.cuh"

namespace milab200 {
using namespace gemv;
namespace {

template <int FMT, int NT>
__global__ void __launch_bounds__(kThreads, 2)
gemv_mma_kernel(const GemvParams p)
{
    using T = FmtTraits<FMT>;
    constexpr bool kIsFp4 = (FMT != kFp8);

    extern __shared__ __align__(16) uint8_t smem_raw[];
    // layout: [xscale: 16 f32][warp max: 8*16 f32][red: tpc*8*NT*128 f32][xs: chunk activations]
    float* s_xscale = reinterpret_cast<float*>(smem_raw);
    float* s_wmax = s_xscale + kMaxTok;
    float* s_red = s_wmax + kWarps * kMaxTok;
    uint4* s_xs = reinterpret_cast<uint4*>(s_red + (size_t)p.tpc * kWarps * NT * 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int M = p.M, K = p.K, N = p.N;
    const int64_t row_bytes = kIsFp4 ? (K >> 1) : K;
    const int groups = p.steps;                     // FP4: one scale per step

    const int tile0 = blockIdx.x * p.tpc;
    const int ntile = min(p.tpc, p.tiles - tile0);

    // ---- zero the reduction buffer --------------------------------------------------------
    for (int i = tid; i < p.tpc * kWarps * NT * 128; i += kThreads) s_red[i] = 0.0f;

    // ---- per-token absmax -> power-of-two scale -------------------------------------------
    {
        float mx[kMaxTok];
#pragma unroll
        for (int m = 0; m < kMaxTok; ++m) mx[m] = 0.0f;
        const int chunks = K >> 3;
#pragma unroll
        for (int m = 0; m < kMaxTok; ++m) {
            if (m < M) {
                const __nv_bfloat16* xr = p.x + (size_t)m * K;
                float a = 0.0f;
                for (int c = tid; c < chunks; c += kThreads) {
                    const uint4 v = ldg_cached_v4(xr + c * 8);
                    const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // compare magnitudes as integers: order-preserving for finite values and
                        // keeps inf/NaN out of the way (handled below)
                        a = fmaxf(a, fabsf(bf16lo(w4[i])));
                        a = fmaxf(a, fabsf(bf16hi(w4[i])));
                    }
                }
                mx[m] = warp_max(a);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < kMaxTok; ++m) s_wmax[warp * kMaxTok + m] = mx[m];
        }
    }
    __syncthreads();
    if (tid < kMaxTok) {
        float a = 0.0f;
        for (int w = 0; w < kWarps; ++w) a = fmaxf(a, s_wmax[w * kMaxTok + tid]);
        // scaled max lands in [2^14, 2^15): e = exponent(a) - 14, clamped so 2^-e and 2^e are normal
        int e = 0;
        if (a > 0.0f && a < __int_as_float(0x7f800000)) {
            e = (int)((__float_as_uint(a) >> 23) & 0xFF) - 127 - 14;
            e = max(-126, min(126, e));
        }
        s_xscale[tid] = __int_as_float((127 + e) << 23);     // 2^e
    }
    __syncthreads();

    // ---- main loop over activation chunks ---------------------------------------------------
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[n][j] = 0.0f;

    for (int step0 = 0; step0 < p.steps; step0 += p.chunk_steps) {
        const int csteps = min(p.chunk_steps, p.steps - step0);
        if (step0 > 0) __syncthreads();             // previous chunk fully consumed

        // stage activations of this chunk: FP16, scaled, in B-fragment order
        {
            const int cpt = csteps * (T::STEP / 8);            // 8-element chunks per token
            for (int idx = tid; idx < cpt * M; idx += kThreads) {
                const int m = idx / cpt, c = idx - m * cpt;
                const int k = c * 8;                             // within the chunk
                const int s = k / T::STEP, within = k - s * T::STEP;
                const int tt = within / T::KT, q = (within - tt * T::KT) >> 3;
                const uint4 v = ldg_cached_v4(p.x + (size_t)m * K + (size_t)step0 * T::STEP + k);
                const float inv = __int_as_float((254 << 23) - __float_as_int(s_xscale[m]));  // 2^-e
                uint4 o;
                o.x = pack_f16x2_rn(bf16lo(v.x) * inv, bf16hi(v.x) * inv);
                o.y = pack_f16x2_rn(bf16lo(v.y) * inv, bf16hi(v.y) * inv);
                o.z = pack_f16x2_rn(bf16lo(v.z) * inv, bf16hi(v.z) * inv);
                o.w = pack_f16x2_rn(bf16lo(v.w) * inv, bf16hi(v.w) * inv);
                s_xs[((s * T::Q + q) * M + m) * 4 + tt] = o;
            }
        }
        __syncthreads();

        // this warp's steps inside the chunk: s = warp, warp + 8, ...
        const int spw = (csteps - warp + kWarps - 1) / kWarps;          // may be 0
        const int nitems = ntile * spw;

        WFrag<FMT> wf[kDepth];
        float sc_lo[kDepth], sc_hi[kDepth];

        auto issue = [&](int item, int slot) {
            const int tl = item / spw, sl = item - tl * spw;
            const int step = step0 + warp + sl * kWarps;
            const int r0 = min((tile0 + tl) * 16 + g, N - 1);
            const int r1 = min((tile0 + tl) * 16 + g + 8, N - 1);
            const uint8_t* plo = p.w + (size_t)r0 * row_bytes + (size_t)step * T::ROWB + t * (T::ROWB / 4);
            const uint8_t* phi = p.w + (size_t)r1 * row_bytes + (size_t)step * T::ROWB + t * (T::ROWB / 4);
            load_w<FMT>(wf[slot], plo, phi);
            if constexpr (kIsFp4) {
                sc_lo[slot] = __ldg(p.scales + (size_t)r0 * groups + step);
                sc_hi[slot] = __ldg(p.scales + (size_t)r1 * groups + step);
            }
        };

#pragma unroll
        for (int d = 0; d < kDepth; ++d)
            if (d < nitems) issue(d, d);

        for (int base = 0; base < nitems; base += kDepth) {
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const int item = base + d;
                if (item < nitems) {
                    const int tl = item / spw, sl = item - tl * spw;
                    const uint4* xs_step = s_xs + (size_t)(warp + sl * kWarps) * T::Q * M * 4;
                    if constexpr (kIsFp4) {
                        float part[NT][4];
#pragma unroll
                        for (int n = 0; n < NT; ++n)
#pragma unroll
                            for (int j = 0; j < 4; ++j) part[n][j] = 0.0f;
                        step_mma<FMT, NT>(part, wf[d], xs_step, M, g, t);
                        const float slo = sc_lo[d], shi = sc_hi[d];
#pragma unroll
                        for (int n = 0; n < NT; ++n) {
                            acc[n][0] = fmaf(part[n][0], slo, acc[n][0]);
                            acc[n][1] = fmaf(part[n][1], slo, acc[n][1]);
                            acc[n][2] = fmaf(part[n][2], shi, acc[n][2]);
                            acc[n][3] = fmaf(part[n][3], shi, acc[n][3]);
                        }
                    } else {
                        step_mma<FMT, NT>(acc, wf[d], xs_step, M, g, t);
                    }
                    if (item + kDepth < nitems) issue(item + kDepth, d);
                    if (sl == spw - 1) {            // tile finished for this chunk: flush
                        float* r = s_red + ((size_t)(tl * kWarps + warp) * NT * 4) * 32 + lane;
#pragma unroll
                        for (int n = 0; n < NT; ++n)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                r[(n * 4 + j) * 32] += acc[n][j];
                                acc[n][j] = 0.0f;
                            }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- cross-warp reduction + epilogue ------------------------------------------------------
    for (int o = tid; o < ntile * NT * 128; o += kThreads) {
        const int tl = o / (NT * 128), r = o - tl * (NT * 128);
        const int n = r >> 7, j = (r >> 5) & 3, ln = r & 31;
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            sum += s_red[((size_t)(tl * kWarps + w) * NT * 4 + n * 4 + j) * 32 + ln];
        const int row = (tile0 + tl) * 16 + (ln >> 2) + ((j >> 1) << 3);
        const int tok = n * 8 + ((ln & 3) << 1) + (j & 1);
        if (row < N && tok < M) {
            float v = sum * s_xscale[tok];
            if constexpr (!kIsFp4) v *= __ldg(p.scales + row);
            if (p.bias) v += __bfloat162float(p.bias[row]);
            p.y[(size_t)tok * N + row] = __float2bfloat16_rn(v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Generic fallback (any K % 8 == 0 the MMA kernel's step size does not divide): one warp per
// output row, FP32 FMA.  Correctness path for odd shapes, not a performance path.
// ------------------------------------------------------------------------------------------
template <bool kIsFp4>
__global__ void __launch_bounds__(256)
gemv_generic_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                    const uint8_t* __restrict__ w, const float* __restrict__ scales,
                    const __nv_bfloat16* __restrict__ bias, int M, int K, int N, int group_size)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    float acc[kMaxTok];
#pragma unroll
    for (int m = 0; m < kMaxTok; ++m) acc[m] = 0.0f;

    for (int k = lane * 8; k < K; k += 32 * 8) {
        float wv[8];
        if constexpr (kIsFp4) {
            const uint32_t word = *reinterpret_cast<const uint32_t*>(w + (size_t)row * (K >> 1) + (k >> 1));
            const float s = scales[(size_t)row * (K / group_size) + k / group_size];
#pragma unroll
            for (int i = 0; i < 8; ++i) wv[i] = e2m1_to_f32((word >> (4 * i)) & 0xFu) * s;
        } else {
            const uint2 two = *reinterpret_cast<const uint2*>(w + (size_t)row * K + k);
            const uint32_t ww[2] = { two.x, two.y };
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint32_t lo, hi;
                cvt_e4m3x4_to_f16x2x2(ww[i], lo, hi);
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                wv[4 * i] = a.x; wv[4 * i + 1] = a.y; wv[4 * i + 2] = b.x; wv[4 * i + 3] = b.y;
            }
        }
#pragma unroll
        for (int m = 0; m < kMaxTok; ++m) {
            if (m < M) {
                const uint4 v = ldg_cached_v4(x + (size_t)m * K + k);
                acc[m] = fmaf(bf16lo(v.x), wv[0], acc[m]); acc[m] = fmaf(bf16hi(v.x), wv[1], acc[m]);
                acc[m] = fmaf(bf16lo(v.y), wv[2], acc[m]); acc[m] = fmaf(bf16hi(v.y), wv[3], acc[m]);
                acc[m] = fmaf(bf16lo(v.z), wv[4], acc[m]); acc[m] = fmaf(bf16hi(v.z), wv[5], acc[m]);
                acc[m] = fmaf(bf16lo(v.w), wv[6], acc[m]); acc[m] = fmaf(bf16hi(v.w), wv[7], acc[m]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < kMaxTok; ++m) {
        if (m < M) {
            float v = warp_sum(acc[m]);
            if (lane == 0) {
                if constexpr (!kIsFp4) v *= scales[row];
                if (bias) v += __bfloat162float(bias[row]);
                y[(size_t)m * N + row] = __float2bfloat16_rn(v);
            }
        }
    }
}

struct DeviceInfo { int sms = 0; int max_smem = 0; bool ok = false; };
const DeviceInfo& device_info()
{
    static thread_local DeviceInfo info[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { static DeviceInfo bad; return bad; }
    if (!info[dev].ok) {
        cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info[dev].max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        info[dev].ok = info[dev].sms > 0;
    }
    return info[dev];
}

template <int FMT, int NT>
int launch_mma(const GemvParams& p, size_t smem, cudaStream_t stream, const char* name)
{
    static thread_local size_t configured[16] = { 0 };
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && smem > configured[dev]) {
        MILAB200_RETURN_IF_CUDA(cudaFuncSetAttribute(gemv_mma_kernel<FMT, NT>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    const int grid = (p.tiles + p.tpc - 1) / p.tpc;
    gemv_mma_kernel<FMT, NT><<<grid, kThreads, smem, stream>>>(p);
    note_launch(name);
    return (int)cudaGetLastError();
}

}  // namespace

template <int FMT>
int try_gemv_flat(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                  int, int, int, cudaStream_t, const char*, const char*, int*);
int try_decode_tc(int fmt, __nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                  int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu);
int try_decode_mx4(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                   int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu);

namespace {

template <int FMT>
int gemv_dispatch(__nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                  const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream,
                  const char* name1, const char* name2, const char* flat1, const char* flat2)
{
    using T = FmtTraits<FMT>;
    const DeviceInfo& di = device_info();
    if (!di.ok) return MILAB200_E_NO_DEVICE;

    // FP4 g = 128, M <= 8: packed-nibble kind::mxf4 kernels (decode_mx4.cu)
    if constexpr (FMT == kFp4G128) {
        int status = 0;
        if (try_decode_mx4(y, x, w, scales, bias, M, K, N, stream, &status, nullptr, 0) == 0) return status;
    }
    // primary path: TMA + tcgen05 stream-K kernel (decode_tc.cu)
    {
        int status = 0;
        if (try_decode_tc(FMT, y, x, w, scales, bias, M, K, N, stream, &status, nullptr, 0) == 0) return status;
    }
    // mma.sync path: persistent row-balanced kernel (activations of all M tokens resident in smem)
    {
        int status = 0;
        if (try_gemv_flat<FMT>(y, x, w, scales, bias, M, K, N, stream, flat1, flat2, &status) == 0)
            return status;
    }

    GemvParams p;
    p.y = y; p.x = x; p.w = w; p.scales = scales; p.bias = bias;
    p.M = M; p.K = K; p.N = N;
    p.tiles = (N + 15) / 16;
    p.steps = K / T::STEP;
    const int NT = (M > 8) ? 2 : 1;

    // row tiles per CTA: amortise the activation staging (M*K elements per CTA) over more rows
    // as M grows, but keep at least ~3 CTAs per SM worth of grid.
    int tpc = (M <= 1) ? 1 : (M <= 2) ? 2 : (M <= 4) ? 2 : (M <= 8) ? 4 : 8;
    while (tpc > 1 && (p.tiles + tpc - 1) / tpc < di.sms * 3) tpc >>= 1;
    p.tpc = tpc;

    const size_t fixed = (size_t)(kMaxTok + kWarps * kMaxTok) * 4 + (size_t)tpc * kWarps * NT * 128 * 4;
    // activation chunk: as many steps as fit in the budget (multiple of 8 so warps stay balanced)
    const size_t budget = 96 * 1024;
    const size_t per_step = (size_t)T::STEP * M * 2;
    int chunk = (int)((budget - fixed) / per_step);
    if (chunk >= p.steps) chunk = p.steps;
    else chunk = max(kWarps, chunk / kWarps * kWarps);
    p.chunk_steps = chunk;
    const size_t smem = fixed + (size_t)chunk * per_step;
    if ((int)smem > di.max_smem) return MILAB200_E_BAD_SHAPE;

    return (NT == 1) ? launch_mma<FMT, 1>(p, smem, stream, name1)
                     : launch_mma<FMT, 2>(p, smem, stream, name2);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// entry points used by abi.cu (M <= 16 per call; callers split larger M)
// ------------------------------------------------------------------------------------------

// route option "decode_generic": every decode call takes the one-warp-per-row kernel — an independent second device
// implementation (plain FP32 FMAs, no tensor cores) the parity tests cross-check the MMA paths against
static std::atomic<bool> g_force_generic{false};
void gemv_set_force_generic(bool on) { g_force_generic.store(on); }
int launch_gemv_generic(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);

int launch_gemv_fp8(void* y, const void* x, const void* w, const float* scales, const void* bias,
                    int M, int K, int N, cudaStream_t stream)
{
    if (!y || !x || !w || !scales || M <= 0 || M > kMaxTok || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (K % 8 != 0) return MILAB200_E_BAD_SHAPE;
    if (g_force_generic.load(std::memory_order_relaxed)) return launch_gemv_generic(y, x, w, scales, bias, M, K, N, 0, stream);
    auto* Y = static_cast<__nv_bfloat16*>(y);
    auto* X = static_cast<const __nv_bfloat16*>(x);
    auto* W = static_cast<const uint8_t*>(w);
    auto* B = static_cast<const __nv_bfloat16*>(bias);
    if (K % FmtTraits<kFp8>::STEP == 0)
        return gemv_dispatch<kFp8>(Y, X, W, scales, B, M, K, N, stream,
                                   "gemv_mma_kernel<fp8,nt1>", "gemv_mma_kernel<fp8,nt2>",
                                   "gemv_flat_kernel<fp8,nt1>", "gemv_flat_kernel<fp8,nt2>");
    gemv_generic_kernel<false><<<(N + 7) / 8, 256, 0, stream>>>(Y, X, W, scales, B, M, K, N, 0);
    note_launch("gemv_generic_kernel<fp8>");
    return (int)cudaGetLastError();
}

int launch_gemv_fp4(void* y, const void* x, const void* w, const float* scales, const void* bias,
                    int M, int K, int N, int group_size, cudaStream_t stream)
{
    if (!y || !x || !w || !scales || M <= 0 || M > kMaxTok || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    if (K % group_size != 0 || K % 8 != 0) return MILAB200_E_BAD_SHAPE;
    if (g_force_generic.load(std::memory_order_relaxed)) return launch_gemv_generic(y, x, w, scales, bias, M, K, N, group_size, stream);
    auto* Y = static_cast<__nv_bfloat16*>(y);
    auto* X = static_cast<const __nv_bfloat16*>(x);
    auto* W = static_cast<const uint8_t*>(w);
    auto* B = static_cast<const __nv_bfloat16*>(bias);
    if (group_size == 128)
        return gemv_dispatch<kFp4G128>(Y, X, W, scales, B, M, K, N, stream,
                                       "gemv_mma_kernel<fp4g128,nt1>", "gemv_mma_kernel<fp4g128,nt2>",
                                       "gemv_flat_kernel<fp4g128,nt1>", "gemv_flat_kernel<fp4g128,nt2>");
    return gemv_dispatch<kFp4G64>(Y, X, W, scales, B, M, K, N, stream,
                                  "gemv_mma_kernel<fp4g64,nt1>", "gemv_mma_kernel<fp4g64,nt2>",
                                  "gemv_flat_kernel<fp4g64,nt1>", "gemv_flat_kernel<fp4g64,nt2>");
}

// reference-semantics generic kernels, exposed for tests (an independent second GPU path)
int launch_gemv_generic(void* y, const void* x, const void* w, const float* scales, const void* bias,
                        int M, int K, int N, int group_size /*0 = fp8*/, cudaStream_t stream)
{
    if (!y || !x || !w || !scales || M <= 0 || M > kMaxTok || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (K % 8 != 0) return MILAB200_E_BAD_SHAPE;
    auto* Y = static_cast<__nv_bfloat16*>(y);
    auto* X = static_cast<const __nv_bfloat16*>(x);
    auto* W = static_cast<const uint8_t*>(w);
    auto* B = static_cast<const __nv_bfloat16*>(bias);
    if (group_size == 0)
        gemv_generic_kernel<false><<<(N + 7) / 8, 256, 0, stream>>>(Y, X, W, scales, B, M, K, N, 0);
    else {
        if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
        if (K % group_size != 0) return MILAB200_E_BAD_SHAPE;
        gemv_generic_kernel<true><<<(N + 7) / 8, 256, 0, stream>>>(Y, X, W, scales, B, M, K, N, group_size);
    }
    note_launch(group_size ? "gemv_generic_kernel<fp4>" : "gemv_generic_kernel<fp8>");
    return (int)cudaGetLastError();
}

}  // namespace milab200
