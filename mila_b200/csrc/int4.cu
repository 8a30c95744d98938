// int4.cu — PerGroupInt4<g> (GPTQ-style W4A16) Linear forward: SURVEY.md §8f rank 4.
//
// Replaces cuda_w4a16_gemm (LIN/Kernels/W4A16Gemm/CudaW4A16Gemm.cuh:73, kernel CudaW4A16Gemm.cu:88-197; call sites
// LIN/CudaLinearOp.ixx:560,786,866):  Y[m,n] = bf16( sum_k X[m,k] * (nibble(n,k) - zero(n,k/g)) * scale(n,k/g) + bias[n] ),
// weights packed two unsigned INT4 per byte (low nibble = even k), one FP32 scale and one INT4 zero point per group of g
// input channels (zero_points == nullptr: symmetric, zero = 8).  The reference has no quantizer for this policy
// (CudaLinearOp.ixx:385-391 throws) — checkpoints arrive pre-quantized — and its kernel is a 16x16 FP32 tile loop.
//
// Unsigned 4-bit integers are not a tcgen05 operand type (kind::f8f6f4 reads a nibble as E2M1, which is not linear in
// the nibble), so this path streams the packed weights with 128-bit loads and feeds mma.sync.m16n8k16 BF16:
//   * (nibble - zero) is made exact in BF16 in two instructions per pair: (w >> 4i) & 0x000F000F | 0x43004300 is the
//     bf16x2 (128 + n_i, 128 + n_{i+4}); one HSUB2.BF16 with (128 + zero) gives the signed integers exactly.  That pairs
//     k = i with k = i + 4 instead of adjacent k — free, because the k slots of an MMA may be permuted as long as the
//     activation fragment uses the same permutation (four PRMTs per 8 activations);
//   * activations are BF16 operands as they are (no scaling, no split); products are exact, accumulation is FP32;
//   * every 128-k (or 64-k) group accumulates in its own MMA accumulators and is folded into the FP32 row sum with the
//     group scale — the factoring the FP4 matvec uses (CudaMatVecBias.Bf16.cu:461-494);
//   * a CTA owns 16 output rows; its 8 warps cut the k range into 8 runs of whole groups, keep 4 groups of weight loads
//     in flight each, and add their partials in warp order through shared memory — same bits every run.
// M <= 16 per launch (tokens are the MMA n dimension); larger M runs in blocks of 16 tokens (the weights of a layer fit
// L2 or stream again: the reference's own kernel reads them M/16 times as well).
#include "common.cuh"

namespace milab200 {
namespace {

constexpr int kWarps = 8, kThreads = kWarps * 32, kRows = 16, kDepth = 4;

struct Int4Params {
    __nv_bfloat16*       y;
    const __nv_bfloat16* x;
    const uint8_t*       w;
    const float*         scales;
    const uint8_t*       zp;
    const __nv_bfloat16* bias;
    int M, K, N, KG;                 // KG = K / G groups
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hsub2_bf16(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// G = quantisation group size (64 or 128), NT = 8-token tiles (1: M <= 8, 2: M <= 16)
template <int G, int NT>
__global__ void __launch_bounds__(kThreads, 2)
w4a16_int4_kernel(const Int4Params p)
{
    constexpr int WPG = G / 32;                       // 32-k chunks (one uint4 per lane-quad member) per group: 4 or 2
    __shared__ float s_part[kWarps][NT][kRows][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = blockIdx.x * kRows + g, row1 = row0 + 8;
    const bool live0 = row0 < p.N, live1 = row1 < p.N;
    const int ra = live0 ? row0 : p.N - 1, rb = live1 ? row1 : p.N - 1;          // clamp: loads stay in bounds, results unused
    // this warp's run of 128-k chunk-quads: a "step" is 128 k (4 lanes x 32 k) whatever G is
    const int steps = p.K / 128;                                                    // K % 128 == 0 on this path
    const int s_begin = warp * steps / kWarps, s_end = (warp + 1) * steps / kWarps;
    const size_t pitch = (size_t)p.K / 2;
    const uint8_t* w0 = p.w + (size_t)ra * pitch + t * 16;
    const uint8_t* w1 = p.w + (size_t)rb * pitch + t * 16;

    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }

    uint4 wq0[kDepth], wq1[kDepth];
    auto issue = [&](int s, int slot) {
        wq0[slot] = ldg_stream_v4(w0 + (size_t)s * 64);
        wq1[slot] = ldg_stream_v4(w1 + (size_t)s * 64);
    };
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
        if (s_begin + d < s_end) issue(s_begin + d, d);

    for (int s = s_begin; s < s_end; ++s) {
        const int slot = (s - s_begin) % kDepth;
        const uint4 q0 = wq0[slot], q1 = wq1[slot];
        if (s + kDepth < s_end) issue(s + kDepth, slot);
        // group of this lane's 32 k: k = s*128 + t*32 .. +31 (32 | G, so the whole chunk shares one scale / zero)
        const int grp = (s * 128 + t * 32) / G;
        const float sc0 = __ldg(p.scales + (size_t)ra * p.KG + grp), sc1 = __ldg(p.scales + (size_t)rb * p.KG + grp);
        uint32_t z0 = 8u, z1 = 8u;
        if (p.zp) {
            const uint32_t b0 = __ldg(p.zp + (size_t)ra * (p.KG / 2) + grp / 2), b1 = __ldg(p.zp + (size_t)rb * (p.KG / 2) + grp / 2);
            z0 = (grp & 1) ? (b0 >> 4) : (b0 & 0xFu);
            z1 = (grp & 1) ? (b1 >> 4) : (b1 & 0xFu);
        }
        const uint32_t zb0 = (0x4300u | z0) * 0x00010001u, zb1 = (0x4300u | z1) * 0x00010001u;   // bf16x2 (128 + zero)
        // activations of this lane's tokens at the same 32 k
        uint4 xa[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int m = j * 8 + g;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                xa[j][c] = make_uint4(0, 0, 0, 0);
                if (m < p.M) xa[j][c] = __ldg(reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K + (size_t)s * 128 + t * 32 + c * 8));
            }
        }
        // a lane's 32 k may straddle nothing, but the four lanes of a quad may sit in different groups when G = 64
        // (t = 0,1 -> group 2s, t = 2,3 -> group 2s + 1): an MMA sums over all four lanes' k, so with G = 64 the scale is
        // applied per lane BEFORE the MMA is impossible — instead each half-quad's k is fed to its own accumulator by
        // zeroing the other half's A operand (two MMA passes).  G = 128: one pass.
        const uint32_t w0w[4] = { q0.x, q0.y, q0.z, q0.w }, w1w[4] = { q1.x, q1.y, q1.z, q1.w };
        constexpr int PASSES = (G == 128) ? 1 : 2;
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
            float d[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) { d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.0f; }
            const bool mine = (PASSES == 1) || ((t >> 1) == pass);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t a[2][4];                       // [row half][pair i]: (n_i, n_{i+4}) - zero, exact bf16x2
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t p0 = ((w0w[c] >> (4 * i)) & 0x000F000Fu) | 0x43004300u;
                    const uint32_t p1 = ((w1w[c] >> (4 * i)) & 0x000F000Fu) | 0x43004300u;
                    a[0][i] = mine ? hsub2_bf16(p0, zb0) : 0u;
                    a[1][i] = mine ? hsub2_bf16(p1, zb1) : 0u;
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const uint4 xv = xa[j][c];
                    const uint32_t x04 = __byte_perm(xv.x, xv.z, 0x5410), x15 = __byte_perm(xv.x, xv.z, 0x7632);
                    const uint32_t x26 = __byte_perm(xv.y, xv.w, 0x5410), x37 = __byte_perm(xv.y, xv.w, 0x7632);
                    mma_bf16_16816(d[j], a[0][0], a[1][0], a[0][1], a[1][1], x04, x15);
                    mma_bf16_16816(d[j], a[0][2], a[1][2], a[0][3], a[1][3], x26, x37);
                }
            }
            // fold this group's exact integer-weighted sums into the FP32 row sums with the group scale.  The scale a lane
            // holds belongs to ITS k chunk; after the MMA every lane of the quad holds sums over the whole quad's k, so the
            // scale must be the (pass-uniform) group's: take it from the quad lane that owns the group.
            const int src = (PASSES == 1) ? (lane & ~3) : ((lane & ~3) | (pass << 1));
            const float g0 = __shfl_sync(0xffffffffu, sc0, src), g1 = __shfl_sync(0xffffffffu, sc1, src);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                acc[j][0] = fmaf(d[j][0], g0, acc[j][0]); acc[j][1] = fmaf(d[j][1], g0, acc[j][1]);
                acc[j][2] = fmaf(d[j][2], g1, acc[j][2]); acc[j][3] = fmaf(d[j][3], g1, acc[j][3]);
            }
        }
    }

    // ---- deterministic cross-warp sum: warp order, then bias, one BF16 rounding ----
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s_part[warp][j][g][2 * t] = acc[j][0];     s_part[warp][j][g][2 * t + 1] = acc[j][1];
        s_part[warp][j][g + 8][2 * t] = acc[j][2]; s_part[warp][j][g + 8][2 * t + 1] = acc[j][3];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < NT * kRows * 8; o += kThreads) {
        const int j = o / (kRows * 8), r = (o / 8) % kRows, c = o % 8;
        const int m = j * 8 + c, row = blockIdx.x * kRows + r;
        if (m < p.M && row < p.N) {
            float v = 0.0f;
#pragma unroll
            for (int wv = 0; wv < kWarps; ++wv) v += s_part[wv][j][r][c];
            if (p.bias) v += __bfloat162float(p.bias[row]);
            p.y[(size_t)m * p.N + row] = __float2bfloat16_rn(v);
        }
    }
}

// any K % group_size == 0 shape the MMA kernel does not take (K % 128 != 0): one warp per output row, plain FP32
__global__ void __launch_bounds__(256)
w4a16_int4_generic_kernel(const Int4Params p, int G)
{
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= p.N) return;
    for (int m = 0; m < p.M; ++m) {
        float acc = 0.0f;
        for (int grp = 0; grp < p.KG; ++grp) {
            float zero = 8.0f;
            if (p.zp) {
                const uint32_t b = p.zp[(size_t)row * (p.KG / 2) + grp / 2];
                zero = (float)((grp & 1) ? (b >> 4) : (b & 0xFu));
            }
            float part = 0.0f;
            for (int k = grp * G + lane; k < (grp + 1) * G; k += 32) {
                const uint32_t byte = p.w[(size_t)row * (p.K / 2) + k / 2];
                const float nib = (float)((k & 1) ? (byte >> 4) : (byte & 0xFu));
                part = fmaf(__bfloat162float(p.x[(size_t)m * p.K + k]), nib - zero, part);
            }
            acc = fmaf(warp_sum(part), p.scales[(size_t)row * p.KG + grp], acc);
        }
        if (lane == 0) p.y[(size_t)m * p.N + row] = __float2bfloat16_rn(acc + (p.bias ? __bfloat162float(p.bias[row]) : 0.0f));
    }
}

}  // namespace

int launch_w4a16_int4(void* out, const void* act, const void* w, const float* scales, const void* zero_points, const void* bias,
                      int M, int K, int N, int group_size, cudaStream_t stream)
{
    if (!out || !act || !w || !scales || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    if (K % group_size != 0 || K % 8 != 0) return MILAB200_E_BAD_SHAPE;
    if (zero_points && (K / group_size) % 2 != 0) return MILAB200_E_BAD_SHAPE;       // two zero points per byte (CudaW4A16Gemm.cu:107)
    Int4Params p;
    p.w = static_cast<const uint8_t*>(w); p.scales = scales; p.zp = static_cast<const uint8_t*>(zero_points);
    p.bias = static_cast<const __nv_bfloat16*>(bias); p.K = K; p.N = N; p.KG = K / group_size;
    const bool mma_ok = (K % 128 == 0) && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0;
    for (int m0 = 0; m0 < M; m0 += 16) {
        const int mb = (M - m0 < 16) ? (M - m0) : 16;
        p.y = static_cast<__nv_bfloat16*>(out) + (size_t)m0 * N;
        p.x = static_cast<const __nv_bfloat16*>(act) + (size_t)m0 * K;
        p.M = mb;
        if (mma_ok) {
            const int grid = (N + kRows - 1) / kRows;
            if (group_size == 128) {
                if (mb <= 8) w4a16_int4_kernel<128, 1><<<grid, kThreads, 0, stream>>>(p);
                else         w4a16_int4_kernel<128, 2><<<grid, kThreads, 0, stream>>>(p);
            } else {
                if (mb <= 8) w4a16_int4_kernel<64, 1><<<grid, kThreads, 0, stream>>>(p);
                else         w4a16_int4_kernel<64, 2><<<grid, kThreads, 0, stream>>>(p);
            }
            note_launch(group_size == 128 ? (mb <= 8 ? "w4a16_int4_kernel<g128,nt1>" : "w4a16_int4_kernel<g128,nt2>")
                                          : (mb <= 8 ? "w4a16_int4_kernel<g64,nt1>" : "w4a16_int4_kernel<g64,nt2>"));
        } else {
            w4a16_int4_generic_kernel<<<(N + 7) / 8, 256, 0, stream>>>(p, group_size);
            note_launch("w4a16_int4_generic_kernel");
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace milab200

extern "C" int milab200_w4a16_gemm(void* out_bf16, const void* act_bf16, const void* weights_packed, const float* scales,
                                   const void* zero_points, const void* bias_bf16, int outer_size, int in_features,
                                   int out_features, int group_size, milab200_stream_t stream)
{
    return milab200::launch_w4a16_int4(out_bf16, act_bf16, weights_packed, scales, zero_points, bias_bf16, outer_size,
                                       in_features, out_features, group_size, static_cast<cudaStream_t>(stream));
}
