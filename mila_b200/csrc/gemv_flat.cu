// gemv_flat.cu — persistent, row-balanced decode kernel (the fast path for M*K*2 <= ~200 KB).
//
// Same per-step machinery as gemv.cu (128-bit streaming loads -> F2FP pair conversion ->
// mma.sync m16n8k16 with the tokens as the n-tile, FP32 accumulate), different work decomposition,
// designed around what the first B200 measurements showed (profiles/r1_perf_shapes_v1_baseline):
// a CTA per 16-row tile with its own activation-staging prologue is latency-bound and loses up to
// 25 % to wave quantisation.  Here:
//   * grid = min(2 x SMs, ceil(N/16)) persistent CTAs; CTA b owns the ROW range
//     [b*N/G, (b+1)*N/G) — balanced to one row, not to a 16-row tile.  A ragged last tile simply
//     skips the loads of rows it does not own, so HBM bytes are balanced across SMs exactly.
//   * inside the CTA the (tile, k-step) space is flattened and cut into 8 equal contiguous warp
//     ranges; a warp streams its range with kDepth steps (1 KB each) in flight and never waits
//     for another warp inside the loop.
//   * the first kDepth weight steps are issued BEFORE the activation prologue, so HBM is busy
//     while the CTA computes token scales and stages FP16 activations (one pass, registers).
//   * tiles cut by a warp boundary are finished deterministically: each warp parks at most two
//     partial 16 x NT*8 tiles in shared memory, and after one __syncthreads the warp that owns
//     the tile's first step adds the later warps' partials in warp order.  (Mila's tests compare
//     two forwards of the same weights with EXPECT_EQ — Linear.Cuda.cpp:744 — so no atomics.)
#include "gemv_common.cuh"

namespace milab200 {
using namespace gemv;
namespace {

constexpr int kXCache = 8;          // uint4 activation chunks a thread keeps between the two prologue phases

struct FlatParams {
    __nv_bfloat16*       y;
    const __nv_bfloat16* x;
    const uint8_t*       w;
    const float*         scales;
    const __nv_bfloat16* bias;
    int M, K, N;
    int steps;                      // K / STEP
};

__device__ __forceinline__ float finite_abs(uint32_t f32bits)
{
    const uint32_t a = f32bits & 0x7FFFFFFFu;
    return (a < 0x7F800000u) ? __uint_as_float(a) : 0.0f;
}

__device__ __forceinline__ float chunk_absmax(const uint4& v)
{
    const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a = fmaxf(a, finite_abs(w4[i] << 16));
        a = fmaxf(a, finite_abs(w4[i] & 0xFFFF0000u));
    }
    return a;
}

template <int FMT, int NT>
__global__ void __launch_bounds__(kThreads, 2)
gemv_flat_kernel(const FlatParams p)
{
    using T = FmtTraits<FMT>;
    constexpr bool kIsFp4 = (FMT != kFp8);
    constexpr int kSlotFloats = NT * 128;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    // [absmax bits: 16 i32][slot tile ids: 8*2 i32][slots: 8*2*NT*128 f32][xs: M*K f16]
    int*   s_absbits = reinterpret_cast<int*>(smem_raw);
    int*   s_slot_tile = s_absbits + kMaxTok;
    float* s_slot = reinterpret_cast<float*>(s_slot_tile + kWarps * 2);
    uint4* s_xs = reinterpret_cast<uint4*>(s_slot + kWarps * 2 * kSlotFloats);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int M = p.M, K = p.K, N = p.N, steps = p.steps;
    const int64_t row_bytes = kIsFp4 ? (K >> 1) : K;

    // ---- this CTA's rows, this warp's flat (tile, step) range ---------------------------------
    const int G = gridDim.x;
    const int r_begin = (int)((int64_t)blockIdx.x * N / G);
    const int r_end = (int)((int64_t)(blockIdx.x + 1) * N / G);
    const int ntiles = (r_end - r_begin + 15) >> 4;
    const int total = ntiles * steps;
    const int a_begin = (int)((int64_t)warp * total / kWarps);
    const int a_end = (int)((int64_t)(warp + 1) * total / kWarps);
    const int count = a_end - a_begin;

    // ---- weight prefetch pipeline ------------------------------------------------------------
    WFrag<FMT> wf[kDepth];
    float sc_lo[kDepth], sc_hi[kDepth];
    int i_tile = (steps > 0) ? a_begin / steps : 0;
    int i_step = a_begin - i_tile * steps;

    auto issue = [&](int slot) {
        const int row_lo = r_begin + i_tile * 16 + g, row_hi = row_lo + 8;
        const size_t off = (size_t)i_step * T::ROWB + t * (T::ROWB / 4);
        if constexpr (FMT == kFp4G64) { wf[slot].lo = make_uint2(0, 0); wf[slot].hi = make_uint2(0, 0); }
        else { wf[slot].lo = make_uint4(0, 0, 0, 0); wf[slot].hi = make_uint4(0, 0, 0, 0); }
        sc_lo[slot] = 0.0f; sc_hi[slot] = 0.0f;
        if (row_lo < r_end) {
            if constexpr (FMT == kFp4G64) wf[slot].lo = ldg_stream_v2(p.w + (size_t)row_lo * row_bytes + off);
            else                          wf[slot].lo = ldg_stream_v4(p.w + (size_t)row_lo * row_bytes + off);
            if constexpr (kIsFp4) sc_lo[slot] = __ldg(p.scales + (size_t)row_lo * steps + i_step);
        }
        if (row_hi < r_end) {
            if constexpr (FMT == kFp4G64) wf[slot].hi = ldg_stream_v2(p.w + (size_t)row_hi * row_bytes + off);
            else                          wf[slot].hi = ldg_stream_v4(p.w + (size_t)row_hi * row_bytes + off);
            if constexpr (kIsFp4) sc_hi[slot] = __ldg(p.scales + (size_t)row_hi * steps + i_step);
        }
        if (++i_step == steps) { i_step = 0; ++i_tile; }
    };

#pragma unroll
    for (int d = 0; d < kDepth; ++d)
        if (d < count) issue(d);

    // ---- prologue: token scales + FP16 activations in B-fragment order -----------------------------
    if (tid < kMaxTok) s_absbits[tid] = 0;
    if (tid < kWarps * 2) s_slot_tile[tid] = -1;
    __syncthreads();

    const int cpt = K >> 3;                         // 8-element chunks per token
    const int nchunks = cpt * M;
    // A warp's 32 consecutive chunks touch at most two tokens when cpt >= 32: reduce each side with
    // shuffles and issue two shared atomics per warp; tiny K takes one atomic per lane.
    auto absmax_accumulate = [&](int c, bool live, float a) {
        const int m = live ? c / cpt : -1;
        if (cpt >= 32) {
            const int m0 = __shfl_sync(0xffffffffu, m, 0);
            const float a0 = warp_max(m == m0 ? a : 0.0f);
            const float a1 = warp_max((m != m0 && m >= 0) ? a : 0.0f);
            if (lane == 0 && m0 >= 0) {
                atomicMax(&s_absbits[m0], __float_as_int(a0));
                if (m0 + 1 < M) atomicMax(&s_absbits[m0 + 1], __float_as_int(a1));
            }
        } else if (live) {
            atomicMax(&s_absbits[m], __float_as_int(a));
        }
    };
    uint4 xc[kXCache];
#pragma unroll
    for (int i = 0; i < kXCache; ++i) {
        const int cw = (tid & ~31) + i * kThreads;            // warp-uniform guard around the shuffles
        if (cw < nchunks) {
            const int c = tid + i * kThreads;
            const bool live = c < nchunks;
            xc[i] = live ? ldg_cached_v4(p.x + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
            absmax_accumulate(c, live, chunk_absmax(xc[i]));
        }
    }
    for (int cw = kXCache * kThreads + (tid & ~31); cw < nchunks; cw += kThreads) {   // rare: M*K > 16384
        const int c = cw + lane;
        const bool live = c < nchunks;
        const uint4 v = live ? ldg_cached_v4(p.x + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
        absmax_accumulate(c, live, chunk_absmax(v));
    }
    __syncthreads();

    auto token_exp = [&](int m) -> int {            // e such that absmax * 2^-e lies in [2^14, 2^15)
        const int bits = s_absbits[m];
        int e = 0;
        if (bits > 0) e = max(-126, min(126, ((bits >> 23) & 0xFF) - 127 - 14));
        return e;
    };
    auto stage_chunk = [&](int c, const uint4& v) {
        const int m = c / cpt, k = (c - m * cpt) * 8;
        const int s = k / T::STEP, within = k - s * T::STEP;
        const int tt = within / T::KT, q = (within - tt * T::KT) >> 3;
        const float inv = __int_as_float((127 - token_exp(m)) << 23);
        uint4 o;
        o.x = pack_f16x2_rn(bf16lo(v.x) * inv, bf16hi(v.x) * inv);
        o.y = pack_f16x2_rn(bf16lo(v.y) * inv, bf16hi(v.y) * inv);
        o.z = pack_f16x2_rn(bf16lo(v.z) * inv, bf16hi(v.z) * inv);
        o.w = pack_f16x2_rn(bf16lo(v.w) * inv, bf16hi(v.w) * inv);
        s_xs[((s * T::Q + q) * M + m) * 4 + tt] = o;
    };
#pragma unroll
    for (int i = 0; i < kXCache; ++i) {
        const int c = tid + i * kThreads;
        if (c < nchunks) stage_chunk(c, xc[i]);
    }
    for (int c = kXCache * kThreads + tid; c < nchunks; c += kThreads)
        stage_chunk(c, ldg_cached_v4(p.x + (size_t)c * 8));
    __syncthreads();

    // ---- epilogue helper: 16 x NT*8 tile values -> y --------------------------------------------
    auto write_tile = [&](const float (&v)[NT][4], int tile) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = r_begin + tile * 16 + g + ((j >> 1) << 3);
                const int tok = n * 8 + (t << 1) + (j & 1);
                if (row < r_end && tok < M) {
                    float r = v[n][j] * __int_as_float((127 + token_exp(tok)) << 23);
                    if constexpr (!kIsFp4) r *= __ldg(p.scales + row);
                    if (p.bias) r += __bfloat162float(p.bias[row]);
                    p.y[(size_t)tok * N + row] = __float2bfloat16_rn(r);
                }
            }
    };

    // ---- main loop ------------------------------------------------------------------------------
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[n][j] = 0.0f;

    int c_tile = (steps > 0) ? a_begin / steps : 0;
    int c_step = a_begin - c_tile * steps;
    int seg_first_step = c_step;            // step at which the current tile segment started
    int nslots = 0;

    for (int base = 0; base < count; base += kDepth) {
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int item = base + d;
            if (item < count) {
                const uint4* xs_step = s_xs + (size_t)c_step * T::Q * M * 4;
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
                if (item + kDepth < count) issue(d);

                const bool tile_done = (c_step == steps - 1);
                if (tile_done || item == count - 1) {
                    if (tile_done && seg_first_step == 0) {
                        write_tile(acc, c_tile);                         // whole tile is ours
                    } else {                                             // cut by a warp boundary
                        float* sl = s_slot + (size_t)(warp * 2 + nslots) * kSlotFloats + lane;
#pragma unroll
                        for (int n = 0; n < NT; ++n)
#pragma unroll
                            for (int j = 0; j < 4; ++j) sl[(n * 4 + j) * 32] = acc[n][j];
                        if (lane == 0)      // bit 30 marks "this segment holds the tile's first step"
                            s_slot_tile[warp * 2 + nslots] = c_tile | (seg_first_step == 0 ? (1 << 30) : 0);
                        ++nslots;
                    }
#pragma unroll
                    for (int n = 0; n < NT; ++n)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[n][j] = 0.0f;
                    seg_first_step = 0;
                }
                if (++c_step == steps) { c_step = 0; ++c_tile; }
            }
        }
    }
    __syncthreads();

    // ---- finish the tiles that were cut: owner = warp holding the tile's first step ----------------
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int tag = s_slot_tile[warp * 2 + s];
        if (tag >= 0 && (tag & (1 << 30))) {
            const int tile = tag & ~(1 << 30);
            float v[NT][4];
            const float* sl = s_slot + (size_t)(warp * 2 + s) * kSlotFloats + lane;
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[n][j] = sl[(n * 4 + j) * 32];
            for (int w2 = warp + 1; w2 < kWarps; ++w2) {
                if (s_slot_tile[w2 * 2] == tile) {                       // later warps: always their slot 0
                    const float* s2 = s_slot + (size_t)(w2 * 2) * kSlotFloats + lane;
#pragma unroll
                    for (int n = 0; n < NT; ++n)
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[n][j] += s2[(n * 4 + j) * 32];
                }
            }
            write_tile(v, tile);
        }
    }
}

struct DevInfo { int sms = 0; int max_smem = 0; bool ok = false; };
const DevInfo& dev_info()
{
    static thread_local DevInfo info[16];
    static DevInfo bad;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return bad;
    if (!info[dev].ok) {
        cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info[dev].max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        info[dev].ok = info[dev].sms > 0;
    }
    return info[dev];
}

template <int FMT, int NT>
int launch_flat(const FlatParams& p, size_t smem, int grid, cudaStream_t stream, const char* name)
{
    static thread_local size_t configured[16] = { 0 };
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && smem > configured[dev]) {
        MILAB200_RETURN_IF_CUDA(cudaFuncSetAttribute(gemv_flat_kernel<FMT, NT>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    gemv_flat_kernel<FMT, NT><<<grid, kThreads, smem, stream>>>(p);
    note_launch(name);
    return (int)cudaGetLastError();
}

}  // namespace

// Returns 1 when the shape is not eligible (caller falls back to the K-chunked kernel), else the
// launch status.
template <int FMT>
int try_gemv_flat(__nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                  const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream,
                  const char* name1, const char* name2, int* status)
{
    using T = FmtTraits<FMT>;
    const DevInfo& di = dev_info();
    if (!di.ok) { *status = MILAB200_E_NO_DEVICE; return 0; }
    const int NT = (M > 8) ? 2 : 1;
    const size_t fixed = (size_t)(kMaxTok + kWarps * 2) * 4 + (size_t)kWarps * 2 * NT * 128 * 4;
    const size_t smem = fixed + (size_t)M * K * 2;
    if (K % T::STEP != 0 || (int)smem > di.max_smem - 1024) return 1;

    FlatParams p;
    p.y = y; p.x = x; p.w = w; p.scales = scales; p.bias = bias;
    p.M = M; p.K = K; p.N = N; p.steps = K / T::STEP;
    // two CTAs per SM when their activations fit side by side, otherwise one
    const int per_sm = (2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
    int grid = di.sms * per_sm;
    const int tiles = (N + 15) / 16;
    if (grid > tiles) grid = tiles;
    *status = (NT == 1) ? launch_flat<FMT, 1>(p, smem, grid, stream, name1)
                        : launch_flat<FMT, 2>(p, smem, grid, stream, name2);
    return 0;
}

template int try_gemv_flat<kFp8>(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*,
                                 const __nv_bfloat16*, int, int, int, cudaStream_t, const char*, const char*, int*);
template int try_gemv_flat<kFp4G128>(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*,
                                     const __nv_bfloat16*, int, int, int, cudaStream_t, const char*, const char*, int*);
template int try_gemv_flat<kFp4G64>(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*,
                                    const __nv_bfloat16*, int, int, int, cudaStream_t, const char*, const char*, int*);

}  // namespace milab200
