// gemv_common.cuh — pieces shared by the decode kernels (gemv.cu: K-chunked variant,
// gemv_flat.cu: persistent row-balanced variant): format traits, weight fragments, and the
// convert + mma.sync step.  See gemv.cu's header comment for the design rationale.
#pragma once
#include "common.cuh"

namespace milab200 {
namespace gemv {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kDepth = 4;        // weight steps in flight per warp
constexpr int kMaxTok = 16;

enum Fmt { kFp8 = 0, kFp4G128 = 1, kFp4G64 = 2 };

// Peer-memory view of a tensor-parallel group for the fused row-parallel all-reduce (tp.cu owns the memory,
// decode_tc.cu's epilogue uses it).  Every rank's exchange buffer holds
//   row_epoch [nmax] u32 (local use only)   and   data [2 parities][world sources][kMaxTok][nmax] {f32 bits, epoch},
// mapped into this process for every rank (CUDA IPC); index [rank] is the local buffer.
constexpr int kTpMaxWorld = 8;
constexpr int kTpMaxTiles = 4096;
struct TpExchange {
    int world = 1, rank = 0, nmax = 0;
    uint2* data[kTpMaxWorld] = {};
    uint32_t* row_epoch = nullptr;         // local [nmax]: calls that have finished this output row so far
};

template <int FMT> struct FmtTraits;
template <> struct FmtTraits<kFp8>     { static constexpr int KT = 16, STEP = 64,  ROWB = 64, Q = 2; };
template <> struct FmtTraits<kFp4G128> { static constexpr int KT = 32, STEP = 128, ROWB = 64, Q = 4; };
template <> struct FmtTraits<kFp4G64>  { static constexpr int KT = 16, STEP = 64,  ROWB = 32, Q = 2; };
// KT   = k elements per thread per step, STEP = 4*KT = k elements per warp-step,
// ROWB = weight bytes per row per step,  Q    = 128-bit activation loads per step per token.

struct GemvParams {
    __nv_bfloat16*       y;        // [M, N]
    const __nv_bfloat16* x;        // [M, K]
    const uint8_t*       w;        // fp8 [N,K] or packed fp4 [N,K/2]
    const float*         scales;   // [N] or [N, K/g]
    const __nv_bfloat16* bias;     // [N] or null
    int M, K, N;
    int tiles;                     // ceil(N/16)
    int tpc;                       // row tiles per CTA
    int steps;                     // K / STEP
    int chunk_steps;               // steps whose activations are staged in smem at a time
};

// Weight bytes one thread holds for one step: rows g and g+8 of the tile.
template <int FMT> struct WFrag { uint4 lo, hi; };
template <> struct WFrag<kFp4G64> { uint2 lo, hi; };

template <int FMT>
__device__ __forceinline__ void load_w(WFrag<FMT>& f, const uint8_t* plo, const uint8_t* phi)
{
    if constexpr (FMT == kFp4G64) { f.lo = ldg_stream_v2(plo); f.hi = ldg_stream_v2(phi); }
    else                          { f.lo = ldg_stream_v4(plo); f.hi = ldg_stream_v4(phi); }
}

// One k-step of one 16-row tile: convert the weight bytes and issue the MMAs against the staged
// activations.  `d` is the accumulator the MMAs add into.
template <int FMT, int NT>
__device__ __forceinline__ void step_mma(float (&d)[NT][4], const WFrag<FMT>& wf,
                                         const uint4* __restrict__ xs_step, int M, int g, int t)
{
    using T = FmtTraits<FMT>;
    // activation fragments: xs_step[(q*M + m)*4 + t]
    const bool tok0 = g < M;
    const bool tok1 = (NT == 2) && (g + 8 < M);

    if constexpr (FMT == kFp8) {
        const uint32_t lo[4] = { wf.lo.x, wf.lo.y, wf.lo.z, wf.lo.w };
        const uint32_t hi[4] = { wf.hi.x, wf.hi.y, wf.hi.z, wf.hi.w };
#pragma unroll
        for (int q = 0; q < T::Q; ++q) {
            uint4 b0 = make_uint4(0, 0, 0, 0), b1 = make_uint4(0, 0, 0, 0);
            if (tok0) b0 = xs_step[(q * M + g) * 4 + t];
            if (NT == 2 && tok1) b1 = xs_step[(q * M + g + 8) * 4 + t];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int i = 2 * q + c;
                uint32_t a0, a2, a1, a3;
                cvt_e4m3x4_to_f16x2x2(lo[i], a0, a2);
                cvt_e4m3x4_to_f16x2x2(hi[i], a1, a3);
                mma_m16n8k16_f16(d[0], a0, a1, a2, a3, c ? b0.z : b0.x, c ? b0.w : b0.y);
                if constexpr (NT == 2)
                    mma_m16n8k16_f16(d[1], a0, a1, a2, a3, c ? b1.z : b1.x, c ? b1.w : b1.y);
            }
        }
    } else {
        constexpr int NW = (FMT == kFp4G128) ? 4 : 2;
        uint32_t lo[NW], hi[NW];
        if constexpr (FMT == kFp4G128) {
            lo[0] = wf.lo.x; lo[1] = wf.lo.y; lo[2] = wf.lo.z; lo[3] = wf.lo.w;
            hi[0] = wf.hi.x; hi[1] = wf.hi.y; hi[2] = wf.hi.z; hi[3] = wf.hi.w;
        } else {
            lo[0] = wf.lo.x; lo[1] = wf.lo.y; hi[0] = wf.hi.x; hi[1] = wf.hi.y;
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) {     // word q of the row == activation chunk q
            uint4 b0 = make_uint4(0, 0, 0, 0), b1 = make_uint4(0, 0, 0, 0);
            if (tok0) b0 = xs_step[(q * M + g) * 4 + t];
            if (NT == 2 && tok1) b1 = xs_step[(q * M + g + 8) * 4 + t];
            uint32_t pl[4], ph[4];
            cvt_e2m1x8_to_f16x2x4(lo[q], pl);
            cvt_e2m1x8_to_f16x2x4(hi[q], ph);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                mma_m16n8k16_f16(d[0], pl[2 * c], ph[2 * c], pl[2 * c + 1], ph[2 * c + 1],
                                 c ? b0.z : b0.x, c ? b0.w : b0.y);
                if constexpr (NT == 2)
                    mma_m16n8k16_f16(d[1], pl[2 * c], ph[2 * c], pl[2 * c + 1], ph[2 * c + 1],
                                     c ? b1.z : b1.x, c ? b1.w : b1.y);
            }
        }
    }
}


}  // namespace gemv
}  // namespace milab200
