// prefill_tc.cu — the B200-native batched (M > 16) Linear forward: a dense TMA -> shared memory ->
// tcgen05.mma (kind::f8f6f4) -> TMEM contraction over the quantized weights AS THEY LIE IN HBM.
//
// Replaces, for M > 16, the reference's FP8 2-phase path (cuda_fp8_dequantize_to_bf16 -> cuBLASLt BF16 GEMM
// -> cuda_add_bias, LIN/CudaLinearOp.ixx:609-643), its FP4 W4A8 4-phase path (cuda_fp4_dequantize_to_fp8 ->
// cuda_quantize_bf16_to_fp8_per_token -> cuBLASLt FP8 GEMM -> cuda_fp8_apply_per_token_scales,
// LIN/CudaLinearOp.ixx:660-714) and the fused slots cuda_w8a16_gemm (K/W8A16Gemm/CudaW8A16Gemm.cu:62-162),
// cuda_fp4a16_gemm (K/W4A16Gemm/CudaW4A16Gemm.cu:88-400), cuda_fp4a16_gemm_wmma (.Wmma.cu:145-333).
//
// Design (same operand trick as decode_tc.cu, sized for the compute-bound regime):
//   * NO weight is ever dequantised or re-expanded.  E4M3 bytes / E2M1 nibbles are the MMA A operand
//     (128 rows x 128 k per pipeline stage, TMA box into 128B-swizzled shared memory; nibbles through the
//     16U4_ALIGN16B tensor-map type).  The reference rewrites the whole weight matrix to scratch on every
//     forward (5 N K bytes of traffic for FP8, CudaW8A16Gemm.cuh:10-14).
//   * the BF16 activations are split EXACTLY into two E4M3 planes by a one-pass pre-kernel
//     (act_split_kernel): with the token's power-of-two scale 2^-e (e from the token's absmax over K),
//     v = x 2^-e in [-256, 256], hi = rn_e4m3(v), lo = rn_e4m3(16 (v - hi)); v = hi + lo/16 for every
//     |v| >= 2^-6, otherwise the absolute error is < 2^-21 of the token maximum.  O(M K) work once per
//     forward instead of O(M N K / 128) dequantisation work inside the GEMM.
//   * a CTA tile is 128 weight rows x 128 tokens; the B operand of the MMA is [hi tokens | lo tokens]
//     = 256 columns, so ONE tcgen05.mma (M = 128, N = 256, K = 32) per 32 k does both planes.
//   * FP8 weights: the per-channel scale factors out of the k sum, so a tile accumulates its whole K in
//     TMEM (2 accumulator buffers x 256 columns = all 512 TMEM columns: the epilogue of tile i overlaps the
//     MMAs of tile i + 1).  Epilogue: y = bf16((D_hi + D_lo/16) 2^e scale[n] + bias[n]).
//   * FP4 weights: every 128-k block (== one PerGroupFp4<128> group) lands in its own TMEM buffer and the
//     8 epilogue warps promote it into FP32 registers with the (row, group) scale — the arithmetic of the
//     reference's matvec (Bf16.cu:461-494), which is also what decode_tc.cu does, so prefill and decode
//     agree to FP32 rounding (the reference pins that relation at 1e-1 row_absmax, Linear.Cuda.cpp:773).
//   * persistent CTAs, static tile schedule with the token tile as the fast index: the CTAs running at
//     any moment share a handful of weight row tiles and every activation tile through L2.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4-11 = epilogue
// (TMEM lane quadrant = warp % 4, token half = (warp - 4) / 4).
#include <cuda.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <unordered_map>

#include "act_split.cuh"
#include "gemv_common.cuh"
#include "glu.cuh"
#include "norm.cuh"
#include "sm100.cuh"

namespace milab200 {
using namespace gemv;
using namespace sm100;
int launch_quantize_bf16_to_fp8_per_token(void* x8, float* sA, const void* x, int M, int K, cudaStream_t st);   // aux.cu
namespace {

constexpr int kRows = 128;                 // weight rows per tile (UMMA M)
constexpr int kTok = 128;                  // tokens per tile; UMMA N = 2 * kTok (hi | lo planes)
constexpr int kBK = 128;                   // k per stage == one FP4 scale group == one 128-byte swizzled row
constexpr int kABytes = kRows * 128;       // 16 KB (FP4 nibbles are unpacked to byte containers by TMA)
constexpr int kBBytes = 2 * kTok * 128;    // 32 KB
constexpr int kAccBufs = 2;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
constexpr int kGluStageBytes = kEpiThreads * 64;    // fused GLU: the 32 BF16 projections per epilogue thread that the peer CTA finishes
template <int CG, bool GLU = false> constexpr size_t pf_smem_bytes()
{
    constexpr int stages = (CG == 2) ? 6 : 4;
    return (size_t)stages * (kABytes + kBBytes / CG) + 8 * (2 * stages + 2 * kAccBufs) + 64 + (GLU ? kGluStageBytes + 128 : 0);
}
static_assert(pf_smem_bytes<2, true>() <= 232448, "exceeds 227 KB of shared memory per CTA");
static_assert(pf_smem_bytes<1>() <= 232448 && pf_smem_bytes<2>() <= 232448, "exceeds 227 KB of shared memory per CTA");

struct PfParams {
    __nv_bfloat16*       y;         // [M, N]
    const float*         xs;        // [Mp] per-token 2^e
    const float*         scales;    // FP8 [N], FP4 [N, KB]
    const __nv_bfloat16* bias;      // [N] or null
    int M, K, N, KB, Mp;
    int tok_tiles, tiles;
    int P, items;                   // k-splits per tile (1 = whole K), tiles * P work items
    float* ws;                      // [items * CG][128 tokens][128 rows] FP32 partial tiles (P > 1)
    int* counters;                  // [tiles * CG] arrival tickets, all zero between launches
    uint32_t a_tx_bytes;
    // activation operand: two exact E4M3 planes of 128 tokens (planes = 2: lo plane `lo_base` = Mp rows below the hi plane,
    // token tiles `tok_stride` = 128 rows apart) or ONE per-token-scaled E4M3 plane of 256 tokens (planes = 1, the
    // reference's W4A8 activation format, CudaFp8Prefill.cu:116-165: the two halves of the MMA's 256 columns are then
    // tokens 0..127 and 128..255 of the tile; lo_base = 128, tok_stride = 256)
    int planes, lo_base, tok_stride;
    // GLU (kernel template GLU, CTA pairs): W is fc_gate_up [2 H, K]; rank 0 of a pair stages 128 gate rows, rank 1 the 128 up
    // rows H below; y is [M, H] = glu(bf16(gate projection), bf16(up projection)) — Gemma.Block.ixx:347-349, Llama.Block.ixx:883
    int glu, H;
    // SP ("summed planes", FP4 weights, CTA pairs, whole-K tiles): tiles of 256 tokens; every k block is two pipeline stages
    // — the hi plane, then the lo = rn(v - hi) plane (act_split_kernel sum_planes) — whose MMAs accumulate into the SAME 256
    // TMEM columns, so a group's accumulator holds 256 tokens instead of 128 x (hi | lo): half the TMEM drain and half the
    // FP32 promotion work per useful flop (the per-group promotion held the tensor pipe at 53 %, profiles/r2j13_ncu_prefill_fp4_*)
};

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(pred));
    return pred != 0;
}

// ---- activation pre-pass -------------------------------------------------------------------------
// One CTA per token row: absmax over K -> e -> two E4M3 planes [2][Mp][K] + xs[m] = 2^e.  Rows m >= M of
// the padded planes are zeroed so that the last token tile reads defined bytes.
__global__ void __launch_bounds__(256)
act_split_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ planes, float* __restrict__ xs,
                 int M, int Mp, int K, const NormArgs norm, const int sum_planes)
{
    const int m = blockIdx.x, tid = threadIdx.x;
    uint8_t* hi_row = planes + (size_t)m * K;
    uint8_t* lo_row = planes + ((size_t)Mp + m) * K;
    const int n8 = K >> 3;
    if (m >= M) {
        for (int i = tid; i < n8; i += 256) {
            reinterpret_cast<uint2*>(hi_row)[i] = make_uint2(0, 0);
            reinterpret_cast<uint2*>(lo_row)[i] = make_uint2(0, 0);
        }
        if (tid == 0) xs[m] = 0.0f;
        return;
    }
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)m * K);
    // fused RMSNorm (norm.cuh): warp 0 computes the token's reciprocal RMS in the reference's reduction order; every
    // thread then normalises its chunks in registers — the split below sees exactly the BF16 values RMSNorm would store
    __shared__ float s_rstd;
    if (norm.on) {
        if (tid < 32) { const float rs = rms_rstd_warp(x + (size_t)m * K, K, norm.eps, tid); if (tid == 0) s_rstd = rs; }
        __syncthreads();
    }
    const float rstd = norm.on ? s_rstd : 1.0f;
    auto load8 = [&](int i) {
        uint4 v = __ldg(xr + i);
        if (norm.on) {
            uint4 w8 = make_uint4(0, 0, 0, 0), b8 = make_uint4(0, 0, 0, 0);
            if (norm.weight) w8 = __ldg(reinterpret_cast<const uint4*>(norm.weight) + i);
            if (norm.bias) b8 = __ldg(reinterpret_cast<const uint4*>(norm.bias) + i);
            v = rms_apply8(v, rstd, w8, b8, norm.weight != nullptr, norm.bias != nullptr, norm.weight_offset);
        }
        return v;
    };
    uint32_t am = 0;
    for (int i = tid; i < n8; i += 256) {
        const uint4 v = load8(i);
        am = __vmaxu2(am, __vmaxu2(__vmaxu2(v.x & 0x7FFF7FFFu, v.y & 0x7FFF7FFFu),
                                   __vmaxu2(v.z & 0x7FFF7FFFu, v.w & 0x7FFF7FFFu)));
    }
    uint32_t a16 = max(am & 0xFFFFu, am >> 16);
    a16 = __reduce_max_sync(0xffffffffu, a16);
    __shared__ uint32_t s_am[8];
    if ((tid & 31) == 0) s_am[tid >> 5] = a16;
    __syncthreads();
    uint32_t amax = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) amax = max(amax, s_am[w]);
    const bool nonfinite = (amax & 0x7F80u) == 0x7F80u;
    amax = min(amax, 0x7F7Fu);
    int e = 0;                                              // absmax * 2^-e in [2^7, 2^8)
    if (amax != 0) e = max(-100, min(100, (int)(amax >> 7) - 127 - 7));
    const float inv = __int_as_float((127 - e) << 23);
    for (int i = tid; i < n8; i += 256) {
        const uint4 v = load8(i);
        uint2 hi, lo;
        if (sum_planes) split_e4m3x8_t<1>(v, inv, hi, lo);      // lo = rn(v - hi): the tensor core adds the planes
        else            split_e4m3x8(v, inv, hi, lo);
        if (nonfinite) poison_nonfinite(v, hi);
        reinterpret_cast<uint2*>(hi_row)[i] = hi;
        reinterpret_cast<uint2*>(lo_row)[i] = lo;
    }
    if (tid == 0) xs[m] = __int_as_float((127 + e) << 23);
}

// ---- the GEMM ------------------------------------------------------------------------------------
// CG = 1: one CTA per 128-row x 128-token tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per
// 256-row x 128-token tile — each CTA stages its own 128 weight rows and ONE activation plane (rank 0 the hi
// plane, rank 1 the lo plane; the pair's MMA reads both), which cuts the L2 -> shared-memory bytes per MMA
// cycle from 96 to 64 (FP8) / 80 to 48 (FP4): the CG = 1 kernel was bound by exactly that stream (ncu:
// 72 B/clk/SM of crossbar reads, tensor pipe 76 %).
template <int FMT, int CG, int SP = 0, bool GLU = false>
__global__ void __launch_bounds__(kThreads, 1)
prefill_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const PfParams p)
{
    constexpr bool kIsFp4 = (FMT != kFp8);
    static_assert(!GLU || (CG == 2 && !kIsFp4), "fused GLU: CTA pairs, FP8 weights (whole-K accumulation: the epilogue has a tile time of slack)");
    // SP = 2: summed planes (PfParams); SP = 1: the same 256-token geometry with ONE per-token-scaled E4M3 plane — the
    // reference's own W4A8 activation format for this weight policy (LIN/CudaLinearOp.ixx:660-714), opt-in and lossy
    static_assert(!SP || (kIsFp4 && CG == 2), "256-token FP4 tiles: CTA pairs only");
    constexpr int kPl = (SP == 2) ? 2 : 1;                       // pipeline stages per k block
    constexpr uint32_t kIdesc = umma_idesc(kIsFp4 ? kFmtE2M1 : kFmtE4M3, kFmtE4M3, kRows * CG, 2 * kTok);
    constexpr uint32_t kAccCols = 2 * kTok;                      // 256 columns per accumulator buffer
    constexpr int kBLocal = kBBytes / CG;                        // activation bytes this CTA stages per k block
    constexpr int kStageB = kABytes + kBLocal;
    constexpr int kNumStages = (CG == 2) ? 6 : 4;
    constexpr int kHalfTok = kTok / 2;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0) __trap();
    const uint32_t bars = base + kNumStages * kStageB;
    auto sA = [&](int s) { return base + (uint32_t)s * kStageB; };
    auto sB = [&](int s) { return base + (uint32_t)s * kStageB + kABytes; };
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kNumStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kNumStages + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kNumStages + kAccBufs + s); };
    uint32_t* g_tmem_base = reinterpret_cast<uint32_t*>(smem_raw + kNumStages * kStageB + 8 * (2 * kNumStages + 2 * kAccBufs));
    // fused GLU: two mbarriers per CTA (peer's half published / my half consumed) and the 16 KB staging area behind them
    const uint32_t glu_full_bar = bars + 8u * (2 * kNumStages + 2 * kAccBufs) + 16u, glu_empty_bar = glu_full_bar + 8u;
    const uint32_t glu_stage = (bars + 8u * (2 * kNumStages + 2 * kAccBufs) + 64u + 127u) & ~127u;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int KB = p.KB;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = (CG == 2) ? (int)cluster_id_x() : (int)blockIdx.x;      // tile-scheduler slot of this CTA (pair)
    const int G = (int)gridDim.x / CG;

    if (tid == 0) {
        for (int s = 0; s < kNumStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        // tempty: one arrival per epilogue warp of every CTA of the pair
        for (int s = 0; s < kAccBufs; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8 * CG); }
        if constexpr (GLU) { mbar_init(glu_full_bar, kEpiThreads); mbar_init(glu_empty_bar, kEpiThreads); }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 2) {
        if constexpr (CG == 2) tmem_alloc_cg2(smem_u32(g_tmem_base), 512);
        else                   tmem_alloc(smem_u32(g_tmem_base), 512);
    }
    tcgen05_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *g_tmem_base;

    if (warp == 0) {
        // ===== TMA producer (every CTA stages its own operands; a pair credits the leader's barrier) =====
        int i = 0;
        for (int item = unit; item < p.items; item += G) {
            const int tile = item / p.P, js = item - tile * p.P;
            const int kbA = (int)((long long)js * KB / p.P), kbB = (int)((long long)(js + 1) * KB / p.P);
            const int rt = tile / p.tok_tiles, tt = tile - rt * p.tok_tiles;
            const int row0 = GLU ? (int)rank * p.H + rt * kRows : (rt * CG + (int)rank) * kRows;
            for (int kbp = kbA * kPl; kbp < kbB * kPl; ++kbp, ++i) {
                const int kb = kbp / kPl, pl = kbp - kb * kPl;   // (SP: plane pl of k block kb; the weights are staged for both)
                const int s = i % kNumStages, ph = (i / kNumStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    if constexpr (SP) {
                        const uint32_t lead_full = mapa_shared(full_bar(s), 0);
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * (p.a_tx_bytes + kBLocal));
                        tma_load_2d_cg2(sA(s), &tmap_w, kb * kBK, row0, lead_full);
                        tma_load_2d_cg2(sB(s), &tmap_x, kb * kBK, pl * p.Mp + tt * p.tok_stride + (int)rank * kTok, lead_full);
                    } else if constexpr (CG == 1) {
                        mbar_arrive_expect_tx(full_bar(s), p.a_tx_bytes + kBBytes);
                        tma_load_2d(sA(s), &tmap_w, kb * kBK, row0, full_bar(s));
                        tma_load_2d(sB(s), &tmap_x, kb * kBK, tt * p.tok_stride, full_bar(s));                            // hi plane
                        tma_load_2d(sB(s) + kTok * 128, &tmap_x, kb * kBK, p.lo_base + tt * p.tok_stride, full_bar(s));   // lo plane
                    } else {
                        const uint32_t lead_full = mapa_shared(full_bar(s), 0);
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * (p.a_tx_bytes + kBLocal));
                        tma_load_2d_cg2(sA(s), &tmap_w, kb * kBK, row0, lead_full);
                        tma_load_2d_cg2(sB(s), &tmap_x, kb * kBK, (int)rank * p.lo_base + tt * p.tok_stride, lead_full);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (the pair's leader CTA only) =====
        if (rank == 0) {
            int i = 0, a = 0;                                    // stage counter, accumulator-buffer use counter
            for (int item = unit; item < p.items; item += G) {
                const int js = item % p.P;
                const int kbA = (int)((long long)js * KB / p.P), kbB = (int)((long long)(js + 1) * KB / p.P);
                for (int kbp = kbA * kPl; kbp < kbB * kPl; ++kbp, ++i) {
                    const int kb = kbp / kPl, pl = kbp - kb * kPl;
                    const int s = i % kNumStages, ph = (i / kNumStages) & 1;
                    const int buf = a % kAccBufs, tph = (a / kAccBufs) & 1;
                    if ((kIsFp4 && pl == 0) || kbp == kbA * kPl) mbar_wait(tempty_bar(buf), tph ^ 1);  // epilogues have drained this buffer
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = umma_desc_k_sw128(sA(s));
                        const uint64_t bdesc = umma_desc_k_sw128(sB(s));
                        const uint32_t d = tmem_base + buf * kAccCols;
                        const bool done = kIsFp4 ? (pl == kPl - 1) : (kb == kbB - 1);
#pragma unroll
                        for (int k = 0; k < kBK / 32; ++k) {    // UMMA K = 32 one-byte containers
                            const uint32_t acc = (kIsFp4 ? pl : kb - kbA) + k > 0;
                            if constexpr (CG == 2) umma_f8f6f4_cg2(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, acc);
                            else                   umma_f8f6f4(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, acc);
                        }
                        if constexpr (CG == 2) {
                            umma_commit_cg2(empty_bar(s), 3);
                            if (done) umma_commit_cg2(tfull_bar(buf), 3);
                        } else {
                            umma_commit(empty_bar(s));
                            if (done) umma_commit(tfull_bar(buf));
                        }
                    }
                    __syncwarp();
                    if (kIsFp4 && pl == kPl - 1) ++a;
                }
                if (!kIsFp4) ++a;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = (weight row r of this CTA's 128, token half h) =====
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        auto release = [&](int buf) {                            // this warp has read everything it needs from `buf`
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0));
                else                   mbar_arrive(tempty_bar(buf));
            }
        };
        // Split-K (P > 1: mid-size M on layers with few row tiles, where whole tiles would leave most SMs idle):
        // every contributor parks its FP32 partial tile ([token][row], unscaled) in the workspace and takes a
        // ticket; the last to arrive adds the P partials in split order — same bits every run — and writes y.
        int* g_flag = reinterpret_cast<int*>(g_tmem_base + 1);
        auto split_arrive = [&](int tile_) -> bool {
            __threadfence();
            bar_sync(1, 256);
            if (tid == 128) {
                const int old = atomicAdd(p.counters + tile_ * CG + (int)rank, 1);
                *g_flag = (old == p.P - 1);
            }
            bar_sync(1, 256);
            const bool last = (*g_flag != 0);
            bar_sync(1, 256);
            if (last) __threadfence();
            return last;
        };
        auto ws_tile = [&](int item_) { return p.ws + ((size_t)item_ * CG + rank) * (kTok * kRows) + r; };
        // Fused GLU (FP8 weights): rank 0 of the pair projects 128 gate rows, rank 1 the 128 up rows H below, with the same
        // (row, token) -> thread mapping.  The activation of a thread's 64 tokens is shared: rank 0 finishes tokens [0, 32),
        // rank 1 tokens [32, 64) — the transcendental math of 64 tokens on one CTA's eight epilogue warps alone cost more than
        // the stand-alone activation kernel (profiles/r2m5_*).  Each rank parks the 32 BF16-rounded projections the peer needs
        // in its own shared memory (64 bytes per thread, 16-byte chunks XOR-swizzled by the row pair: conflict-free for a
        // quarter-warp), releases them with a cluster-scope arrive on the peer's `full` barrier, pulls the peer's half over
        // distributed shared memory and frees the peer's staging area with an arrive on the peer's `empty` barrier.
        // The arithmetic is the two-kernel sequence's: both projections rounded to BF16 exactly as the unfused Linear stores
        // them, activation in FP32 with the reference's expressions (glu.cuh).
        auto pack_bf16 = [](float lo, float hi) {
            return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
        };
        uint32_t glu_round = 0;
        const uint32_t glu_row = glu_stage + (uint32_t)(h * kRows + r) * 64u;
        auto glu_exchange = [&](const uint32_t (&give)[16], uint32_t (&take)[16]) {
            const uint32_t peer = rank ^ 1u;
            mbar_wait_acquire_cluster(glu_empty_bar, (glu_round & 1u) ^ 1u);      // the peer has read my previous round
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                             :: "r"(glu_row + (uint32_t)((c ^ ((r >> 1) & 3)) << 4)), "r"(give[4 * c]), "r"(give[4 * c + 1]), "r"(give[4 * c + 2]),
                                "r"(give[4 * c + 3]) : "memory");
            mbar_arrive_release_cluster(mapa_shared(glu_full_bar, peer));
            mbar_wait_acquire_cluster(glu_full_bar, glu_round & 1u);               // the peer's half is in ITS shared memory
            const uint32_t src = mapa_shared(glu_row, peer);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("ld.shared::cluster.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(take[4 * c]), "=r"(take[4 * c + 1]), "=r"(take[4 * c + 2]), "=r"(take[4 * c + 3])
                             : "r"(src + (uint32_t)((c ^ ((r >> 1) & 3)) << 4)) : "memory");
            mbar_arrive_release_cluster(mapa_shared(glu_empty_bar, peer));
            ++glu_round;
        };
        // mine: this thread's 64 projections (packed BF16 pairs, pair i = tokens tbase + 2 i, + 1)
        auto glu_finish = [&](const uint32_t (&mine)[32], int tbase, int col_, bool row_ok_) {
            uint32_t give[16], take[16], keep[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { give[i] = mine[rank == 0 ? 16 + i : i]; keep[i] = mine[rank == 0 ? i : 16 + i]; }
            glu_exchange(give, take);
            const int tb2 = tbase + (rank == 0 ? 0 : 32);
            // straight-line evaluation of the 32 elements (one branch per element — the activation kind, the reciprocal's
            // slow-path guard, the store predicate — serialised their MUFU latencies: 7.5 us per tile instead of < 1),
            // then the stores
            uint32_t o2[16];
            if (p.glu == kGluGegluTanh) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t g2 = (rank == 0) ? keep[i] : take[i], u2 = (rank == 0) ? take[i] : keep[i];
                    o2[i] = pack_bf16(gelu_tanh_fwd(bf16lo(g2)) * bf16lo(u2), gelu_tanh_fwd(bf16hi(g2)) * bf16hi(u2));
                }
            } else {
                bool slow = false;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t g2 = (rank == 0) ? keep[i] : take[i], u2 = (rank == 0) ? take[i] : keep[i];
                    o2[i] = pack_bf16(silu_fwd_fast(bf16lo(g2), slow) * bf16lo(u2), silu_fwd_fast(bf16hi(g2), slow) * bf16hi(u2));
                }
                if (__any_sync(0xffffffffu, slow)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t g2 = (rank == 0) ? keep[i] : take[i], u2 = (rank == 0) ? take[i] : keep[i];
                        o2[i] = pack_bf16(silu_fwd(bf16lo(g2)) * bf16lo(u2), silu_fwd(bf16hi(g2)) * bf16hi(u2));
                    }
                }
            }
            __nv_bfloat16* yp = p.y + (size_t)tb2 * p.H + col_;
            if (row_ok_ && tb2 + 31 < p.M) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    yp[(size_t)(2 * i) * p.H]     = __ushort_as_bfloat16((unsigned short)(o2[i] & 0xFFFFu));
                    yp[(size_t)(2 * i + 1) * p.H] = __ushort_as_bfloat16((unsigned short)(o2[i] >> 16));
                }
            } else if (row_ok_) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (tb2 + 2 * i < p.M)     yp[(size_t)(2 * i) * p.H]     = __ushort_as_bfloat16((unsigned short)(o2[i] & 0xFFFFu));
                    if (tb2 + 2 * i + 1 < p.M) yp[(size_t)(2 * i + 1) * p.H] = __ushort_as_bfloat16((unsigned short)(o2[i] >> 16));
                }
            }
        };
        int a = 0;
        for (int item = unit; item < p.items; item += G) {
            const int tile = item / p.P, js = item - tile * p.P;
            const int kbA = (int)((long long)js * KB / p.P), kbB = (int)((long long)(js + 1) * KB / p.P);
            const int rt = tile / p.tok_tiles, tt = tile - rt * p.tok_tiles;
            // GLU: `col` is the output column (gate row index), `row` the weight row this CTA projects (gate or up)
            const int col = GLU ? rt * kRows + r : (rt * CG + (int)rank) * kRows + r;
            const int row = GLU ? (int)rank * p.H + col : col;
            const bool row_ok = GLU ? (col < p.H) : (row < p.N);
            const int t0 = tt * kTok + h * kHalfTok;
            float bv = 0.0f;
            if (row_ok && p.bias) bv = __bfloat162float(p.bias[row]);

            if constexpr (!kIsFp4) {
                const int buf = a % kAccBufs, tph = (a / kAccBufs) & 1;
                const float rs = row_ok ? __ldg(p.scales + row) : 0.0f;
                mbar_wait(tfull_bar(buf), tph);
                tcgen05_fence_after();
                uint32_t glu_mine[GLU ? 32 : 1];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t dh[32], dl[32];
                    const uint32_t ta = tmem_base + lane_base + buf * kAccCols + h * kHalfTok + c * 32;
                    tmem_ld_32x32b_x32(ta, dh);
                    tmem_ld_32x32b_x32(ta + kTok, dl);
                    tmem_ld_wait();
                    if (c == 1) release(buf);
                    if constexpr (GLU) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const int t = t0 + c * 32 + j;
                            const float d0 = fmaf(__uint_as_float(dl[j]), 0.0625f, __uint_as_float(dh[j]));
                            const float d1 = fmaf(__uint_as_float(dl[j + 1]), 0.0625f, __uint_as_float(dh[j + 1]));
                            glu_mine[c * 16 + j / 2] = pack_bf16(fmaf(d0 * __ldg(p.xs + t), rs, bv), fmaf(d1 * __ldg(p.xs + t + 1), rs, bv));
                        }
                    } else
                    if (p.P > 1) {
                        float* wp = ws_tile(item) + (size_t)(h * kHalfTok + c * 32) * kRows;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            __stcg(wp + j * kRows, fmaf(__uint_as_float(dl[j]), 0.0625f, __uint_as_float(dh[j])));
                    } else if (p.planes == 1) {
                        // one activation plane: the two column halves are two different tokens, each with its own scale
                        // (y = acc * sA[m] * scale[n] + bias, the reference's cuda_fp8_apply_per_token_scales)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int ta_ = tt * p.tok_stride + h * kHalfTok + c * 32 + j, tb_ = ta_ + kTok;
                            if (row_ok && ta_ < p.M)
                                p.y[(size_t)ta_ * p.N + row] = __float2bfloat16_rn(fmaf(__uint_as_float(dh[j]) * __ldg(p.xs + ta_), rs, bv));
                            if (row_ok && tb_ < p.M)
                                p.y[(size_t)tb_ * p.N + row] = __float2bfloat16_rn(fmaf(__uint_as_float(dl[j]) * __ldg(p.xs + tb_), rs, bv));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int t = t0 + c * 32 + j;
                            if (row_ok && t < p.M) {
                                const float dv = fmaf(__uint_as_float(dl[j]), 0.0625f, __uint_as_float(dh[j]));
                                p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(fmaf(dv * __ldg(p.xs + t), rs, bv));
                            }
                        }
                    }
                }
                ++a;
                if constexpr (GLU) glu_finish(glu_mine, t0, col, row_ok);
                if (!GLU && p.P > 1 && split_arrive(tile)) {
                    // 32 independent loads in flight per split (a dependent chain of L2 round trips otherwise)
#pragma unroll 1
                    for (int jt0 = 0; jt0 < kHalfTok; jt0 += 32) {
                        float dv[32];
#pragma unroll
                        for (int u = 0; u < 32; ++u) dv[u] = 0.0f;
                        for (int jj = 0; jj < p.P; ++jj) {
                            const float* src = ws_tile(tile * p.P + jj) + (size_t)(h * kHalfTok + jt0) * kRows;
#pragma unroll
                            for (int u = 0; u < 32; ++u) dv[u] += __ldcg(src + u * kRows);
                        }
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            const int t = t0 + jt0 + u;
                            if (row_ok && t < p.M)
                                p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(fmaf(dv[u] * __ldg(p.xs + t), rs, bv));
                        }
                    }
                    if (tid == 128) p.counters[tile * CG + (int)rank] = 0;
                }
            } else if constexpr (SP) {
                // summed planes: the group's accumulator IS W (hi + lo) for 256 tokens; this thread promotes row r, tokens
                // tt * 256 + h * 128 + [0, 128) — 128 running sums in registers, 8 chunks of 16 columns per group
                float acc[kTok];
#pragma unroll
                for (int j = 0; j < kTok; ++j) acc[j] = 0.0f;
                const float* sp = p.scales + (size_t)(row_ok ? row : 0) * KB;
                float cur[4], nxt[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) cur[u] = (row_ok && kbA + u < kbB) ? __ldg(sp + kbA + u) : 0.0f;
                for (int kb0 = kbA; kb0 < kbB; kb0 += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) nxt[u] = (row_ok && kb0 + 4 + u < kbB) ? __ldg(sp + kb0 + 4 + u) : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (kb0 + u < kbB) {
                            const int buf = a % kAccBufs, tph = (a / kAccBufs) & 1;
                            const float wsc = cur[u];
                            const uint32_t ta = tmem_base + lane_base + buf * kAccCols + h * kTok;
                            mbar_wait(tfull_bar(buf), tph);
                            tcgen05_fence_after();
                            uint32_t dv[2][16];
                            tmem_ld_32x32b_x16(ta, dv[0]);
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                tmem_ld_wait();
                                if (c < 7) tmem_ld_32x32b_x16(ta + (c + 1) * 16, dv[(c + 1) & 1]);
#pragma unroll
                                for (int j = 0; j < 16; j += 2)
                                    fma_f32x2(acc[c * 16 + j], acc[c * 16 + j + 1], __uint_as_float(dv[c & 1][j]), __uint_as_float(dv[c & 1][j + 1]),
                                              wsc, wsc, acc[c * 16 + j], acc[c * 16 + j + 1]);
                            }
                            release(buf);
                            ++a;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
                }
                if (row_ok) {
                    const int tb = tt * p.tok_stride + h * kTok;
#pragma unroll
                    for (int j = 0; j < kTok; ++j) {
                        const int t = tb + j;
                        if (t < p.M) p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(fmaf(acc[j], __ldg(p.xs + t), bv));
                    }
                }
            } else {
                float acc[kHalfTok];
#pragma unroll
                for (int j = 0; j < kHalfTok; ++j) acc[j] = 0.0f;
                const float* sp = p.scales + (size_t)(row_ok ? row : 0) * KB;
                float cur[4], nxt[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) cur[u] = (row_ok && kbA + u < kbB) ? __ldg(sp + kbA + u) : 0.0f;
                for (int kb0 = kbA; kb0 < kbB; kb0 += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) nxt[u] = (row_ok && kb0 + 4 + u < kbB) ? __ldg(sp + kb0 + 4 + u) : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (kb0 + u < kbB) {
                            const int buf = a % kAccBufs, tph = (a / kAccBufs) & 1;
                            const float wsc = cur[u];
                            const uint32_t ta = tmem_base + lane_base + buf * kAccCols + h * kHalfTok;
                            mbar_wait(tfull_bar(buf), tph);
                            tcgen05_fence_after();
                            // 4 chunks of 16 columns, loads of chunk c + 1 in flight under the FMAs of chunk c
                            uint32_t dh[2][16], dl[2][16];
                            tmem_ld_32x32b_x16(ta, dh[0]);
                            tmem_ld_32x32b_x16(ta + kTok, dl[0]);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                tmem_ld_wait();
                                if (c < 3) {
                                    tmem_ld_32x32b_x16(ta + (c + 1) * 16, dh[(c + 1) & 1]);
                                    tmem_ld_32x32b_x16(ta + kTok + (c + 1) * 16, dl[(c + 1) & 1]);
                                }
#pragma unroll
                                for (int j = 0; j < 16; j += 2) {
                                    float v0, v1;
                                    fma_f32x2(v0, v1, __uint_as_float(dl[c & 1][j]), __uint_as_float(dl[c & 1][j + 1]), 0.0625f, 0.0625f,
                                              __uint_as_float(dh[c & 1][j]), __uint_as_float(dh[c & 1][j + 1]));
                                    fma_f32x2(acc[c * 16 + j], acc[c * 16 + j + 1], v0, v1, wsc, wsc, acc[c * 16 + j], acc[c * 16 + j + 1]);
                                }
                            }
                            release(buf);
                            ++a;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
                }
                if (p.P > 1) {
                    float* wp = ws_tile(item) + (size_t)(h * kHalfTok) * kRows;
#pragma unroll
                    for (int j = 0; j < kHalfTok; ++j) __stcg(wp + j * kRows, acc[j]);
                    if (split_arrive(tile)) {
#pragma unroll
                        for (int j = 0; j < kHalfTok; ++j) acc[j] = 0.0f;
                        for (int jj = 0; jj < p.P; ++jj) {
                            const float* src = ws_tile(tile * p.P + jj) + (size_t)(h * kHalfTok) * kRows;
#pragma unroll
                            for (int j = 0; j < kHalfTok; ++j) acc[j] += __ldcg(src + j * kRows);
                        }
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < kHalfTok; ++j) {
                                const int t = t0 + j;
                                if (t < p.M) p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(fmaf(acc[j], __ldg(p.xs + t), bv));
                            }
                        }
                        if (tid == 128) p.counters[tile * CG + (int)rank] = 0;
                    }
                } else if (row_ok) {
#pragma unroll
                    for (int j = 0; j < kHalfTok; ++j) {
                        const int t = t0 + j;
                        if (t < p.M) p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(fmaf(acc[j], __ldg(p.xs + t), bv));
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, 512);
        else                   tmem_dealloc(tmem_base, 512);
    }
}

// =================================================================================================
// host side
// =================================================================================================
struct MapKey {
    const void* ptr; long long rows; int K, fmt;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && K == o.K && fmt == o.fmt; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= ((size_t)k.rows << 32) ^ ((size_t)k.K << 2) ^ (size_t)k.fmt;
        return h;
    }
};

// [rows, K elements] K-major, box = 128 k x 128 rows, 128B swizzle.  fmt: kFp8 = one byte per element
// (weights or activation planes), else packed nibbles through the 16U4_ALIGN16B type.
int tile_tensor_map(const void* ptr, long long rows, int K, int fmt, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ ptr, rows, K, fmt };
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return MILAB200_E_NO_DEVICE;
    const bool fp4 = (fmt != kFp8);
    const cuuint64_t dims[2] = { (cuuint64_t)K, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)(fp4 ? K / 2 : K) };
    const cuuint32_t box[2] = { (cuuint32_t)kBK, (cuuint32_t)kRows };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = enc(out, fp4 ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                           const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MILAB200_E_BAD_SHAPE;
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

// Library-owned activation workspace of one device: the two E4M3 planes + per-token scales of the
// forward in flight.  Grow-only; growing is an allocation, hence illegal during stream capture
// (milab200_reserve_prefill() sizes it beforehand).  A superseded buffer is RETIRED, never freed: CUDA graphs captured
// earlier have its address baked into their act_split_kernel arguments and tensor maps, and replaying them after a
// later, larger forward must stay valid (buffers grow geometrically, so the retired ones add up to less than the
// live one).  One forward at a time per device, like the
// reference's single-stream ExecutionContext scratch (CudaExecutionContext.ixx:164-270).
struct PfDevice {
    bool checked = false, ok = false;
    int sms = 0;
    uint8_t* planes = nullptr;
    size_t capacity = 0;            // bytes
    std::vector<void*> retired;     // superseded planes buffers, kept alive for graphs captured against them
    float* split_ws = nullptr;      // sms x [128][128] FP32 partial tiles (split-K launches are one wave)
    int* split_counters = nullptr;
};
PfDevice g_pf[16];
std::mutex g_pf_mu;

size_t ws_bytes_for(int Mp, int K) { return (size_t)2 * Mp * K + (size_t)Mp * sizeof(float); }

PfDevice* pf_device(size_t need, cudaStream_t stream)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    PfDevice& d = g_pf[dev];
    std::lock_guard<std::mutex> lk(g_pf_mu);
    if (!d.checked) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        d.ok = (major == 10 && d.sms > 0 && encode_tiled_fn() != nullptr);
        d.checked = true;
        if (d.ok) {
            // fixed-size split-K workspace: allocated with the first use of the device (never inside a capture:
            // the planes buffer is grown under the same rule right below)
            cudaStreamCaptureStatus cs0 = cudaStreamCaptureStatusNone;
            if (stream && cudaStreamIsCapturing(stream, &cs0) != cudaSuccess) { cudaGetLastError(); cs0 = cudaStreamCaptureStatusActive; }
            if (cs0 != cudaStreamCaptureStatusNone) { d.checked = false; return nullptr; }
            const size_t wsb = (size_t)d.sms * kTok * kRows * sizeof(float), ctb = 4096 * sizeof(int);
            if (cudaMalloc(&d.split_ws, wsb) != cudaSuccess || cudaMalloc(&d.split_counters, ctb) != cudaSuccess ||
                cudaMemset(d.split_counters, 0, ctb) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
                cudaGetLastError();
                d.ok = false;
            }
        }
    }
    if (!d.ok) return nullptr;
    if (need > d.capacity) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (stream && cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cs != cudaStreamCaptureStatusNone) return nullptr;
        // the old buffer may still be read by an enqueued forward
        if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return nullptr; }
        const size_t cap = (need + need / 4 > 2 * d.capacity) ? need + need / 4 : 2 * d.capacity;
        uint8_t* fresh = nullptr;
        if (cudaMalloc(&fresh, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (d.planes) d.retired.push_back(d.planes);
        d.planes = fresh;
        d.capacity = cap;
    }
    return &d;
}

template <int FMT, int CG, int SP = 0, bool GLU = false>
int launch_pf(const CUtensorMap& tw, const CUtensorMap& tx, const PfParams& p, int units, cudaStream_t stream, const char* name)
{
    static std::atomic<bool> configured[16];
    constexpr size_t smem = pf_smem_bytes<CG, GLU>();
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !configured[dev].load()) {
        MILAB200_RETURN_IF_CUDA(cudaFuncSetAttribute(prefill_tc_kernel<FMT, CG, SP, GLU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)smem));
        configured[dev].store(true);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units * CG); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = (CG == 2) ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, prefill_tc_kernel<FMT, CG, SP, GLU>, tw, tx, p);
    if (e != cudaSuccess) return (int)e;
    note_launch(name);
    return 0;
}

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
std::atomic<bool> g_pf_enabled{ env_int("MILAB200_PREFILL_TC", 1) != 0 };
// activation planes of the batched FP8-weight path: 2 = exact split (default, the conforming path), 1 = ONE per-token E4M3
// plane — the reference's own lossy W4A8 activation format (CudaFp8Prefill.cu:116-165, gate 1e-1 row_absmax,
// Linear.Cuda.cpp:773): half the MMA work per useful flop.  Opt-in.
std::atomic<int> g_pf_planes{ env_int("MILAB200_PREFILL_ACT_PLANES", 2) };
std::atomic<int> g_pf_cg{ env_int("MILAB200_PREFILL_CG", 2) };
// FP4 weights, batched: summed planes (PfParams, kernel SP) — 0 off (128-token tiles, D_hi + D_lo / 16 in the epilogue),
// 1 where the 256-token geometry fills the CTA pairs without a k split (default), 2 whenever M >= 256 and N >= 256
std::atomic<int> g_pf_fp4_sum{ env_int("MILAB200_PREFILL_FP4_SUM", 1) };
// gate|up Linear + GLU for M > 16: 1 = activation fused into the batched kernel's epilogue (default), 0 = Linear + activation kernel
std::atomic<int> g_pf_glu{ env_int("MILAB200_PREFILL_GLU", 1) };

}  // namespace

// Returns 1 when the shape / device is not eligible (the caller takes the token-blocked decode
// kernels), else 0 with the launch status in *status.
int try_prefill_tc_norm(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                        const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status, const NormArgs* norm);
int try_prefill_tc(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                   const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status)
{
    return try_prefill_tc_norm(fmt, y, x, w, scales, bias, M, K, N, stream, status, nullptr);
}

static int prefill_tc_impl(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                           const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status, const NormArgs* norm, int glu);
int try_prefill_tc_norm(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                        const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status, const NormArgs* norm)
{
    return prefill_tc_impl(fmt, y, x, w, scales, bias, M, K, N, stream, status, norm, 0);
}
// Gate|up Linear [2 H, K] with the gated activation in the epilogue (optionally behind the fused RMSNorm): y is [M, H].
// Returns 1 when the fused kernel does not take the call (the caller runs the Linear and the activation kernel).
int try_prefill_tc_glu(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                       const __nv_bfloat16* bias, int M, int K, int H, int glu_kind, cudaStream_t stream, int* status, const NormArgs* norm)
{
    if (glu_kind != kGluGegluTanh && glu_kind != kGluSwiglu) return 1;
    return prefill_tc_impl(fmt, y, x, w, scales, bias, M, K, 2 * H, stream, status, norm, glu_kind);
}

static int prefill_tc_impl(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                           const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status, const NormArgs* norm, int glu)
{
    if (!g_pf_enabled.load(std::memory_order_relaxed)) return 1;
    if (glu && !g_pf_glu.load(std::memory_order_relaxed)) return 1;
    if (fmt != kFp8 && fmt != kFp4G128) return 1;
    if (M < 1 || K % kBK != 0 || K < kBK) return 1;
    if ((reinterpret_cast<uintptr_t>(w) & 31) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
    // one-plane (FP8-rate) mode: FP8 weights, or FP4 weights on CTA pairs (the 256-token kernel variant SP = 1); no fused
    // norm, enough tokens for 256-token tiles
    const bool fp4_pairs = (fmt == kFp4G128 && g_pf_cg.load(std::memory_order_relaxed) == 2 && N >= 2 * kRows);
    const bool one_plane = (g_pf_planes.load(std::memory_order_relaxed) == 1 && (fmt == kFp8 || fp4_pairs) && !norm && M >= 2 * kTok);
    // fused GLU: CTA pairs (rank 0 gate rows, rank 1 up rows), exact activation planes, whole-K tiles
    const int Hglu = N / 2;
    // (FP8 weights only: with per-group FP32 promotion the FP4 epilogue is on the kernel's critical path and the fused
    // activation measured slower than the stand-alone kernel, 760 vs 420 us on the Gemma shape at M = 2048)
    if (glu && (fmt != kFp8 || one_plane || g_pf_cg.load(std::memory_order_relaxed) != 2)) return 1;
    // summed planes (FP4 weights): 256-token tiles on CTA pairs, whole-K items only — chosen when the 256-token geometry
    // fills the CTA pairs without a k split (the mid-size-M split-K regime keeps the 128-token tiles)
    const int sp_on = g_pf_fp4_sum.load(std::memory_order_relaxed);
    bool sum_planes = false;
    if (sp_on && fp4_pairs && !one_plane && M >= 2 * kTok) {
        int sms = 0, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
            const long long tiles256 = (long long)(((glu ? 2 * Hglu : N) + 2 * kRows - 1) / (2 * kRows)) * ((M + 2 * kTok - 1) / (2 * kTok));
            sum_planes = (sp_on >= 2) || tiles256 * 4 > (long long)(sms / 2) * 3;
        }
    }
    const bool tok256 = one_plane || sum_planes;
    const int Mp = tok256 ? (M + 2 * kTok - 1) / (2 * kTok) * (2 * kTok) : (M + kTok - 1) / kTok * kTok;
    PfDevice* d = pf_device(ws_bytes_for(Mp, K), stream);
    if (!d) return 1;
    if (glu) {
        // the fused activation takes whole-K tiles only: where the tile count calls for the k split below (mid-size M on few
        // row tiles) the caller runs the Linear and the stand-alone activation kernel
        const long long tiles = (long long)((Hglu + kRows - 1) / kRows) * (Mp / (tok256 ? 2 * kTok : kTok));
        if (!tok256 && tiles * 4 <= (long long)(d->sms / 2) * 3 && env_int("MILAB200_PREFILL_SPLITK", 1) && K / kBK >= 8) return 1;
    }

    uint8_t* planes = d->planes;
    float* xs = reinterpret_cast<float*>(planes + (size_t)2 * Mp * K);
    CUtensorMap tw, tx;
    if (tile_tensor_map(w, N, K, fmt, &tw) != 0) return 1;
    if (tile_tensor_map(planes, one_plane ? (long long)Mp : 2LL * Mp, K, kFp8, &tx) != 0) return 1;

    cudaError_t e;
    if (one_plane) {
        // the reference's per-token quantizer, bit for bit (aux.cu); rows M..Mp of the last tile must read as zero
        const int rc = launch_quantize_bf16_to_fp8_per_token(planes, xs, x, M, K, stream);
        if (rc != 0) { *status = rc; return 0; }
        if (Mp > M) e = cudaMemsetAsync(planes + (size_t)M * K, 0, (size_t)(Mp - M) * K, stream); else e = cudaSuccess;
    } else {
        act_split_kernel<<<Mp, 256, 0, stream>>>(x, planes, xs, M, Mp, K, norm ? *norm : NormArgs(), sum_planes ? 1 : 0);
        e = cudaGetLastError();
        note_launch("act_split_kernel");
    }
    if (e != cudaSuccess) { *status = (int)e; return 0; }

    PfParams p;
    p.y = y; p.xs = xs; p.scales = scales; p.bias = bias;
    p.M = M; p.K = K; p.N = N; p.KB = K / kBK; p.Mp = Mp;
    p.planes = one_plane ? 1 : 2;
    p.lo_base = one_plane ? kTok : Mp;
    p.tok_stride = tok256 ? 2 * kTok : kTok;
    p.tok_tiles = Mp / p.tok_stride;
    static const int fp4_tx = env_int("MILAB200_FP4_TX_BYTES", kRows * kBK / 2);
    p.a_tx_bytes = (fmt == kFp8) ? (uint32_t)kABytes : (uint32_t)fp4_tx;
    // CTA pairs (256-row tiles) whenever there are at least two 128-row tiles; a single tile runs unpaired
    const int row_tiles = glu ? 2 * ((Hglu + kRows - 1) / kRows) : (N + kRows - 1) / kRows;    // (GLU: gate tile + up tile per pair)
    const int cg = (g_pf_cg.load(std::memory_order_relaxed) == 2 && row_tiles >= 2) ? 2 : 1;
    p.glu = glu; p.H = Hglu;
    p.tiles = ((row_tiles + cg - 1) / cg) * p.tok_tiles;
    const int slots = d->sms / cg;
    // split K when whole tiles would leave most SMs idle (mid-size M on layers with few row tiles): one wave of
    // tiles x P items, every split at least 4 k blocks long
    static const int split_on = env_int("MILAB200_PREFILL_SPLITK", 1);
    p.P = 1;
    if (split_on && !tok256 && p.tiles * 4 <= slots * 3) {
        p.P = slots / p.tiles;
        if (p.P > p.KB / 4) p.P = p.KB / 4;
        if (p.P > 8) p.P = 8;
        if (p.P < 1) p.P = 1;
        if (p.tiles * cg > 4096) p.P = 1;
    }
    if (glu) p.P = 1;                      // (eligibility was settled before the activation pre-pass)
    p.items = p.tiles * p.P;
    p.ws = d->split_ws; p.counters = d->split_counters;
    const int units = p.items < slots ? p.items : slots;
    if (glu)
        *status = launch_pf<kFp8, 2, 0, true>(tw, tx, p, units, stream, "prefill_tc_kernel<fp8,cta_pair,glu>");
    else if (cg == 2)
        *status = (fmt == kFp8) ? launch_pf<kFp8, 2>(tw, tx, p, units, stream, one_plane ? "prefill_tc_kernel<fp8,cta_pair,a8>" : "prefill_tc_kernel<fp8,cta_pair>")
                : one_plane     ? launch_pf<kFp4G128, 2, 1>(tw, tx, p, units, stream, "prefill_tc_kernel<fp4g128,cta_pair,a8>")
                : sum_planes    ? launch_pf<kFp4G128, 2, 2>(tw, tx, p, units, stream, "prefill_tc_kernel<fp4g128,cta_pair,sum>")
                                : launch_pf<kFp4G128, 2>(tw, tx, p, units, stream, "prefill_tc_kernel<fp4g128,cta_pair>");
    else
        *status = (fmt == kFp8) ? launch_pf<kFp8, 1>(tw, tx, p, units, stream, one_plane ? "prefill_tc_kernel<fp8,a8>" : "prefill_tc_kernel<fp8>")
                                : launch_pf<kFp4G128, 1>(tw, tx, p, units, stream, "prefill_tc_kernel<fp4g128>");
    return 0;
}

void prefill_tc_set_enabled(bool on) { g_pf_enabled.store(on); }
void prefill_tc_set_planes(int n) { g_pf_planes.store(n == 1 ? 1 : 2); }
void prefill_tc_set_cta_group(int cg) { g_pf_cg.store(cg == 1 ? 1 : 2); }
void prefill_tc_set_fp4_sum(int v) { g_pf_fp4_sum.store(v < 0 ? 0 : (v > 2 ? 2 : v)); }
void prefill_tc_set_glu(int v) { g_pf_glu.store(v != 0); }

int prefill_tc_reserve(int max_tokens, int max_in_features)
{
    if (max_tokens <= 0 || max_in_features <= 0) return MILAB200_E_INVALID_ARGUMENT;
    const int Mp = (max_tokens + 2 * kTok - 1) / (2 * kTok) * (2 * kTok);      // (256-token tiles of the one-plane / summed-planes modes)
    return pf_device(ws_bytes_for(Mp, max_in_features), nullptr) ? 0 : MILAB200_E_NO_DEVICE;
}

}  // namespace milab200
