// decode_tc.cu — the B200-native decode (M = 1..16) Linear forward: TMA -> shared memory ->
// tcgen05.mma (kind::f8f6f4) -> TMEM -> FP32 promotion.
//
// Replaces cuda_matvec_decode_bf16_qfp8 / _qfp4 (LIN/Kernels/MatVec/CudaMatVecBias.Bf16.cu:198,
// :271, :376) and serves 2 <= M <= 16 of the batched slots (CudaW8A16Gemm.cu:62, CudaW4A16Gemm.cu:88).
//
// The kernel is HBM-bound, so the design goal is that NO SM instruction touches a weight:
//   * the quantized weights are the A operand of the MMA exactly as they lie in HBM.  A TMA tensor
//     map moves a 128-row x 128-k tile per unit into a 128B-swizzled shared-memory stage: E4M3 bytes
//     as they are, E2M1 nibbles through the 16U4_ALIGN16B tensor-map type, which unpacks 16 nibbles
//     into the 16-byte container tcgen05 expects for 4-bit operands.  tcgen05.mma kind::f8f6f4 takes
//     E4M3 or E2M1 for A independently of B.
//   * the BF16 activations are the B operand.  kind::f8f6f4 wants <= 8-bit types, so each activation
//     is split EXACTLY into two E4M3 numbers: with the block's power-of-two scale 2^-e (e from the
//     token's absmax inside this 128-k block) v = x*2^-e lies in [-256, 256]; hi = rn_e4m3(v),
//     lo = rn_e4m3(16*(v - hi)), and x*2^-e = hi + lo/16 whenever |v| >= 2^-6 (an 8-bit significand
//     is two 4-bit significands); below that the absolute error is < 2^-21 of the block maximum,
//     far under FP32 accumulation noise.  hi and lo are separate MMA columns (N = 16 for <= 8
//     tokens, 32 for <= 16), recombined in FP32.
//   * every 128-k block accumulates into a fresh TMEM slot (128 lanes = 128 weight rows, N columns),
//     and the epilogue warps promote it into FP32 registers:  acc += (D_hi + D_lo/16) * 2^e * s,
//     s = the FP4 group scale of (row, block) — one block == one PerGroupFp4<128> group — or 1 for
//     FP8, whose per-channel scale multiplies once at the end (the factoring of Bf16.cu:249,:461-494).
//     Products are exact in the tensor core; nothing is rounded before FP32.
//   * work decomposition: tiles x P work items (whole 128-row tiles whenever there are enough of them, else the
//     k range of a tile cut into P equal runs so that one wave of items covers the SMs), one persistent CTA
//     per SM.  The P partials of a cut tile meet through distributed shared memory — the P CTAs launch as one
//     thread-block cluster and the leader pulls the others' FP32 partials with ld.shared::cluster — and are
//     added in split order: same bits every run (Mila's tests compare two forwards with EXPECT_EQ,
//     Linear.Cuda.cpp:744).  A stream-K variant (units cut into one equal range per SM) exists for unbalanced
//     two-wave shapes at M > 8; a workspace + atomic-ticket fix-up serves it and launches whose clusters cannot
//     be co-resident.
//   * row-parallel tensor parallelism finishes its all-reduce in this epilogue over NVLink peer memory (tp.cu),
//     and a gate|up Linear can apply its GeGLU / SwiGLU here (glu.cu).
//   * 9..16 tokens: the split of the activations is done ONCE per forward by act_presplit_kernel (below) instead
//     of by every CTA's converter warps, whose latency chains set the unit cadence at that width; the TMA producer
//     bulk-copies the ready-made operand image and the block scales into the stage (bit-identical results).
//   * programmatic dependent launch: weights never depend on the previous kernel, so the TMA
//     producer starts streaming them before griddepcontrol.wait; only the activation converter and
//     the epilogue wait for the previous kernel's results.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4-7 = epilogue (TMEM lanes
// 32*(w-4) .. +31), 8-15 = activation converters (idle when the activations arrive pre-split).
#include <cuda.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "act_split.cuh"
#include "gemv_common.cuh"
#include "glu.cuh"
#include "norm.cuh"
#include "sm100.cuh"

namespace milab200 {
using namespace gemv;
using namespace sm100;
namespace {

constexpr int kTileRows = 128;          // UMMA M
constexpr int kBlockK = 128;            // k elements per group (one swizzled 128-byte row) == one FP4 scale group
constexpr int kXsRing = 32;             // activation-scale ring, one entry per group (>= groups * (stages + tmem units + 2))
constexpr int kScDepth = 16;            // FP4 group scales: two batches of 8 scalars in a cp.async ring per epilogue thread
constexpr int kABytes = kTileRows * 128;            // shared bytes of one group of weights
constexpr int kWsRegions = 4;
constexpr int kMaxSplitItems = 1024;     // workspace slots per region (items of launches with P > 1)
constexpr int kMaxTiles = 4096;
constexpr int kWsSlotFloats = kMaxTok * kTileRows;  // one partial tile: [token][row]

struct TcParams {
    __nv_bfloat16*       y;
    const __nv_bfloat16* x;
    const float*         scales;
    const __nv_bfloat16* bias;
    float*               ws;            // [items][16][128] partial tiles (P > 1 only)
    int*                 counters;      // [tiles] arrival tickets, all zero between launches
    int M, K, N;
    int KB;                             // K / 128 groups
    int KBU;                            // ceil(KB / groups-per-unit) unit blocks
    int R;                              // weight rows per tile (<= 128): the TMA box height.  The MMA always spans 128 rows;
                                        //   rows R..127 of a stage hold stale bytes and their TMEM lanes are never read
                                        //   (row i of D depends on row i of A only)
    int tiles;                          // ceil(N / R)
    int P;                              // k-splits per row tile
    int items;                          // tiles * P work items
    uint32_t a_tx_bytes;                // mbarrier transaction bytes of one weight tile
    int early_ld;                       // griddepcontrol.launch_dependents before (1) or after (0) the set-up
    int cl;                             // split-K partials meet through distributed shared memory (cluster of P CTAs)
    int glu, H;                         // fused gate|up -> GLU epilogue: kind (glu.cuh) and hidden width; then N = 2 H,
                                        // tiles = H / 128 logical tiles and KBU counts the units of BOTH halves
    TpExchange tp;                      // row-parallel shard: partial rows are summed across ranks in the epilogue
    int ps;                             // activations were split by act_presplit_kernel: the producer bulk-copies the
    const uint8_t* xp;                  //   plane image [KBU * groups][NCOLS rows x 128 B, swizzled] and the block
    const float*   xps;                 //   scales [KBU * groups][kMaxTok] instead of the converter warps
    NormArgs norm;                      // RMSNorm folded into the activation path (norm.cuh): x is normalised in registers
    long long* prof;                    // bring-up only: CTA 0 records per-unit role timestamps [unit][16]
};

__device__ __forceinline__ long long globaltimer_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_PROF_CTA(slot)                                                                   \
    do { if constexpr (PROF) { if (p.prof) p.prof[1024 + blockIdx.x * 4 + (slot)] = globaltimer_ns(); } } while (0)
#define TC_PROF(slot)                                                                       \
    do { if constexpr (PROF) { if (p.prof && blockIdx.x == 0 && i < 64) p.prof[i * 16 + (slot)] = clock64() - t_start; } } while (0)

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(pred));
    return pred != 0;
}

template <int NCOLS> struct TcShape {
    static constexpr int HALF = NCOLS / 2;                       // token capacity
    static constexpr int kConvWarps = 8;                         // activation-converter warps (round-robin over units): with 4,
                                                                 // one warp per group of EVERY unit, the converter (1900 cycles at M = 8)
                                                                 // set the unit cadence (2600 vs 1800 cycles at M = 1, tools/tc_timeline.py)
    static constexpr int kThreads = (8 + kConvWarps) * 32;
    static constexpr int kBBytes = NCOLS * 128;
    // groups per unit: a pipeline stage holds 128 rows x (128 * kGroups) k.  Every role pays a fixed
    // latency per unit (barrier round trips, a lone warp's dependent issue), so units are as large as
    // shared memory allows: 512 k for <= 8 tokens, 256 k for <= 16.
    static constexpr int kGroups = (NCOLS == 16) ? 4 : 2;
    static constexpr int kStages = (NCOLS == 16) ? 3 : 5;        // shared-memory ring depth (72 / 40 KB per stage)
    static constexpr int kTmemUnits = 8 / kGroups;               // accumulator ring depth in units (kGroups accumulators each)
    static constexpr int kScBatch = 8 / kGroups;                 // FP4 group scales are fetched 8 scalars at a time
    static constexpr size_t kSmem = (size_t)kStages * kGroups * (kABytes + kBBytes) + kXsRing * kMaxTok * 4 +
                                    8 * (2 * kStages + 2 * kTmemUnits) + 128 + kScDepth * kTileRows * 4;
    static_assert(kXsRing >= kGroups * (kStages + kTmemUnits + 2), "activation-scale ring too short");
};

// Work decomposition: the N x K weight matrix is cut into `items` = tiles x P work items, item
// (tile, j) = 128 rows x the j-th of P equal runs of 256-k unit blocks.  CTA c takes items c, c + G, ...
// P == 1 (every shape with >= ~100 row tiles): an item is a whole row tile and its rows are written
// directly.  P > 1 (few row tiles, long K): the P partial tiles of a row tile meet in a workspace and
// the last contributor to arrive (atomic ticket) adds them in j order — same bits every run.
struct Cursor {
    // Two decompositions of the tiles x KBU units of a launch, chosen on the host:
    //  * p.P > 0 (items): tiles x P work items, item (tile, j) = the j-th of P equal runs of a tile's units;
    //    CTA c takes items c, c + G, ...
    //  * p.P == 0 (stream-K): the units, tile-major, are cut into G equal contiguous ranges, one per CTA, so every
    //    SM streams the same number of bytes whatever N and K are; a range crosses tile boundaries, the segments
    //    of a tile cut by a range boundary meet in the workspace exactly like the P partials of an item split.
    int it, tile, ub, ub_end;       // items: item index / stream-K: current unit, ub_end = end of this CTA's range
    bool sk;
    __device__ __forceinline__ void load(const TcParams& p)
    {
        if (it < p.items) {
            tile = it / p.P;
            const int j = it - tile * p.P;
            ub = (int)((long long)j * p.KBU / p.P);
            ub_end = (int)((long long)(j + 1) * p.KBU / p.P);
        }
    }
    __device__ __forceinline__ void start(int cta, const TcParams& p)
    {
        sk = (p.P == 0);
        if (sk) {
            const long long T = (long long)p.tiles * p.KBU;
            it = (int)(cta * T / gridDim.x);
            ub_end = (int)((cta + 1) * T / gridDim.x);
            tile = it / p.KBU;
            ub = it - tile * p.KBU;
        } else {
            it = cta; load(p);
        }
    }
    __device__ __forceinline__ bool valid(const TcParams& p) const { return sk ? (it < ub_end) : (it < p.items); }
    __device__ __forceinline__ bool item_end(const TcParams& p) const
    {
        return sk ? (ub == p.KBU - 1 || it == ub_end - 1) : (ub == ub_end - 1);
    }
    __device__ __forceinline__ void next(const TcParams& p, int G)
    {
        if (sk) {
            ++it;
            if (++ub == p.KBU) { ub = 0; ++tile; }
        } else if (++ub == ub_end) { it += G; load(p); }
    }
};

// stream-K bookkeeping: first unit of CTA c's range, and the CTA whose range holds unit u
__device__ __forceinline__ int sk_start(int c, long long T, int G) { return (int)(c * T / G); }
__device__ __forceinline__ int sk_owner(int u, long long T, int G) { return (int)((((long long)u + 1) * G - 1) / T); }

template <int FMT, int NCOLS, bool PROF>
__global__ void __launch_bounds__(TcShape<NCOLS>::kThreads, 1)
decode_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const TcParams p)
{
    using Shape = TcShape<NCOLS>;
    constexpr bool kIsFp4 = (FMT != kFp8);
    constexpr int HALF = Shape::HALF;
    constexpr int kBBytes = Shape::kBBytes;               // one group of activations
    constexpr int NCW = Shape::kConvWarps;
    constexpr int kStages = Shape::kStages;
    constexpr int kGroups = Shape::kGroups, kTmemUnits = Shape::kTmemUnits, kScBatch = Shape::kScBatch;
    constexpr int kAStage = kGroups * kABytes, kBStage = kGroups * kBBytes;
    constexpr uint32_t kIdesc = umma_idesc(kIsFp4 ? kFmtE2M1 : kFmtE4M3, kFmtE4M3, kTileRows, NCOLS);
    constexpr uint32_t kTmemCols = kTmemUnits * kGroups * NCOLS;      // 128 or 256: a power of two

    extern __shared__ __align__(1024) uint8_t smem_raw[];        // 128B-swizzled tiles need 1024-byte alignment
    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0) __trap();
    uint8_t* gen_base = smem_raw;
    const uint32_t sA = base;
    const uint32_t sB = sA + kStages * kAStage;
    uint8_t* gB = gen_base + kStages * kAStage;
    float* g_xs = reinterpret_cast<float*>(gB + kStages * kBStage);
    const uint32_t bars = sB + kStages * kBStage + kXsRing * kMaxTok * 4;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + kTmemUnits + s); };
    uint8_t* g_misc = gB + kStages * kBStage + kXsRing * kMaxTok * 4 + 8 * (2 * kStages + 2 * kTmemUnits);
    uint32_t* g_tmem_base = reinterpret_cast<uint32_t*>(g_misc);
    int* g_flag = reinterpret_cast<int*>(g_misc + 4);
    float* g_rstd = reinterpret_cast<float*>(g_misc + 64);      // [kMaxTok] reciprocal RMS per token (fused RMSNorm only)
    float* g_scraw = reinterpret_cast<float*>(g_misc + 128);    // [kScDepth][128] (FP4 only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, KB = p.KB;
    if (tid == 0) TC_PROF_CTA(0);

    // ---- one-time setup -------------------------------------------------------------------------
    // The TMA producer (warp 0) needs nothing but the mbarriers: it initialises them and starts streaming
    // weights at once; it only ARRIVES at the set-up barrier the other warps wait on (TMEM allocation, zeroed
    // activation stages), so the first weight bytes are in flight while the rest of the CTA sets up.
    if (p.early_ld) griddep_launch_dependents();        // the next kernel may start its weight prefetch
    if (warp == 0) {
        if (lane == 0) {
            // full: the producer's expect_tx arrival + one arrival per group from the converter warp that owns it
            // (pre-split activations: the producer's weight arrival + its activation arrival)
            for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), p.ps ? 2 : 1 + kGroups); mbar_init(empty_bar(s), 1); }
            for (int s = 0; s < kTmemUnits; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 128); }
            if (p.cl) { mbar_init(smem_u32(g_misc + 16), 128 * (p.P - 1)); mbar_init(smem_u32(g_misc + 24), 1); }
            fence_mbar_init();
            tma_prefetch_desc(&tmap_w);
        }
        __syncwarp();
        asm volatile("bar.arrive 2, %0;" :: "n"(Shape::kThreads) : "memory");
    } else {
        // unused token rows of the activation stages must read as zero; scale slots must be finite.  (Pre-split
        // activations: the bulk copies fill whole stages — dead rows are zero in the image — and may already be landing,
        // since the producer does not wait for this set-up: the stages must NOT be touched here.)
        if (!p.ps)
            for (int i = tid - 32; i < kStages * kBStage / 16; i += Shape::kThreads - 32)
                reinterpret_cast<uint4*>(gB)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid - 32; i < kScDepth * kTileRows; i += Shape::kThreads - 32) g_scraw[i] = 0.0f;
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc(smem_u32(g_tmem_base), kTmemCols);
        tcgen05_fence_before();
        bar_sync(2, Shape::kThreads);
        tcgen05_fence_after();
    }
    const uint32_t tmem_base = (warp == 0) ? 0u : *g_tmem_base;
    if (!p.early_ld) griddep_launch_dependents();

    const long long t_start = (PROF && p.prof) ? clock64() : 0;
    (void)t_start;
    if (tid == 0) TC_PROF_CTA(1);

    Cursor cur;
    cur.start(blockIdx.x, p);
    // fused GLU: a logical tile streams its gate rows' units (ub < KBH), then its up rows' (tile + H/128)
    const int KBH = p.glu ? p.KBU / 2 : p.KBU;
    auto kbu_of = [&](int ub) { return (p.glu && ub >= KBH) ? ub - KBH : ub; };
    auto prow_of = [&](int tile_, int ub) { return tile_ * p.R + ((p.glu && ub >= KBH) ? p.H : 0); };

    if (warp == 0) {
        // ===== TMA producer (whole warp converged, one elected lane issues): weights do not depend
        //       on the previous kernel, so this role never executes griddepcontrol.wait =====
        const uint64_t policy = l2_policy_evict_first();
        // Pre-split activations (p.ps): the plane image and the block scales of a unit are two more bulk copies into
        // the unit's stage / scale-ring slots.  They ARE the previous kernel's output, so they are issued behind
        // griddepcontrol.wait — after the weight loads of the first kStages units are already in flight.
        Cursor cb = cur;
        int ib = 0, i = 0;
        bool waited = false;
        auto issue_b_upto = [&](int last) {
            if (!waited) { griddep_wait(); waited = true; }
            for (; ib <= last; ++ib, cb.next(p, G)) {
                if (elect_one()) {
                    const int sb = ib % kStages;
                    const size_t kb0 = (size_t)kbu_of(cb.ub) * kGroups;
                    mbar_arrive_expect_tx(full_bar(sb), kGroups * (kBBytes + kMaxTok * 4));
                    bulk_load_1d(sB + sb * kBStage, p.xp + kb0 * kBBytes, kGroups * kBBytes, full_bar(sb));
                    bulk_load_1d(smem_u32(g_xs) + ((ib * kGroups) % kXsRing) * (kMaxTok * 4), p.xps + kb0 * kMaxTok,
                                 kGroups * kMaxTok * 4, full_bar(sb));
                }
                __syncwarp();
            }
        };
        for (; cur.valid(p); ++i, cur.next(p, G)) {
            const int s = i % kStages, ph = (i / kStages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
                TC_PROF(0);
                mbar_arrive_expect_tx(full_bar(s), kGroups * p.a_tx_bytes);
#pragma unroll
                for (int g = 0; g < kGroups; ++g)               // a group past the end of K is zero-filled by TMA
                    tma_load_2d_hint(sA + s * kAStage + g * kABytes, &tmap_w, (kbu_of(cur.ub) * kGroups + g) * kBlockK,
                                     prow_of(cur.tile, cur.ub), full_bar(s), policy);
                TC_PROF(1);
            }
            __syncwarp();
            if (p.ps && i + 1 >= kStages) issue_b_upto(i);
        }
        if (p.ps && ib < i) issue_b_upto(i - 1);                 // fewer than kStages units in this CTA
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues) =====
        for (int i = 0; cur.valid(p); ++i, cur.next(p, G)) {
            const int s = i % kStages, ph = (i / kStages) & 1;
            const int slot = i % kTmemUnits, tph = (i / kTmemUnits) & 1;
            mbar_wait(tempty_bar(slot), tph ^ 1);
            mbar_wait(full_bar(s), ph);
            tcgen05_fence_after();
            if (elect_one()) {
                TC_PROF(7);
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {             // every 128-k group accumulates into its own columns
                    const uint64_t adesc = umma_desc_k_sw128(sA + s * kAStage + g * kABytes);
                    const uint64_t bdesc = umma_desc_k_sw128(sB + s * kBStage + g * kBBytes);
                    const uint32_t d = tmem_base + (slot * kGroups + g) * NCOLS;
#pragma unroll
                    for (int k = 0; k < kBlockK / 32; ++k)      // UMMA K = 32 eight-bit containers = 32 bytes
                        umma_f8f6f4(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, k > 0);
                }
                TC_PROF(13);
                umma_commit(empty_bar(s));                      // stage reusable once the MMAs have read it
                umma_commit(tfull_bar(slot));                   // accumulators ready for the epilogue
                TC_PROF(8);
            }
            __syncwarp();
        }
    } else if (warp >= 8 && p.ps) {
        // pre-split activations: nothing to convert
    } else if (warp >= 8) {
        // ===== activation converters: BF16 -> two E4M3 planes in swizzled K-major rows.  Converter warp
        //       cw owns group cw % kGroups of the units i == cw / kGroups (mod NCW / kGroups) of this CTA, so NCW groups
        //       are converted concurrently and the latency of one conversion (loads, shuffles, cvt chains)
        //       is off the critical path.  Lane (tsub, seg8) handles, for j = 0 .. HALF/2-1, the 8
        //       activations of token 2j + tsub at k = 8*seg8 .. +7; a token's block absmax is a 16-lane
        //       shuffle reduction. =====
        const int cw = warp - 8;
        const int g = cw % kGroups, ustride = NCW / kGroups, ufirst = cw / kGroups;
        const int seg8 = lane & 15, tsub = lane >> 4;
        constexpr int CH = HALF / 2;
        griddep_wait();                                         // x is the previous kernel's output
        if (p.norm.on) {
            // fused RMSNorm: converter warp cw computes the reciprocal RMS of tokens cw, cw + 8 in the reference's own
            // reduction order (norm.cuh); the normalised BF16 activations then exist only in registers
            for (int m = cw; m < p.M; m += NCW) {
                const float rs = rms_rstd_select(p.norm, p.x + (size_t)m * p.K, p.K, lane);
                if (lane == 0) g_rstd[m] = rs;
            }
            asm volatile("bar.sync 3, %0;" :: "n"(NCW * 32) : "memory");
        }
        uint4 nxt[CH];
        uint4 nw8 = make_uint4(0, 0, 0, 0), nb8 = make_uint4(0, 0, 0, 0);      // norm weight / bias of this lane's 8 k
        auto x_load = [&](int ub) {
            const int kb = ub * kGroups + g;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int m = 2 * j + tsub;
                nxt[j] = make_uint4(0, 0, 0, 0);
                if (2 * j < p.M && m < p.M && kb < KB)
                    nxt[j] = __ldcg(reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K + (size_t)kb * kBlockK + seg8 * 8));
            }
            if (p.norm.on && kb < KB) {
                if (p.norm.weight) nw8 = __ldg(reinterpret_cast<const uint4*>(p.norm.weight + (size_t)kb * kBlockK + seg8 * 8));
                if (p.norm.bias) nb8 = __ldg(reinterpret_cast<const uint4*>(p.norm.bias + (size_t)kb * kBlockK + seg8 * 8));
            }
        };
        for (int q = 0; q < ufirst && cur.valid(p); ++q) cur.next(p, G);
        Cursor pre = cur;                                        // runs `ustride` units ahead: this warp's next unit
        if (pre.valid(p)) x_load(kbu_of(pre.ub));
        for (int i = ufirst; cur.valid(p); i += ustride) {
            const int s = i % kStages, ph = (i / kStages) & 1;
            uint4 cx[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) cx[j] = nxt[j];
            if (p.norm.on) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int m = 2 * j + tsub;
                    if (2 * j < p.M && m < p.M)
                        cx[j] = rms_apply8(cx[j], g_rstd[m], nw8, nb8, p.norm.weight != nullptr, p.norm.bias != nullptr, p.norm.weight_offset);
                }
            }
#pragma unroll 1
            for (int q = 0; q < ustride && pre.valid(p); ++q) pre.next(p, G);
            if (pre.valid(p)) x_load(kbu_of(pre.ub));            // register prefetch of this warp's next unit
            if (lane == 0) TC_PROF(2 + (g & 1));
            mbar_wait(empty_bar(s), ph ^ 1);                     // stage free (its previous MMAs retired)
            uint8_t* bstage = gB + s * kBStage + g * kBBytes;
            float* xs_slot = g_xs + ((i * kGroups + g) % kXsRing) * kMaxTok;
            // block absmax per token: packed 16-bit unsigned max over the BF16 magnitudes of 16 lanes x 8 values.
            // All chunks go through each shuffle level together so the shuffle latencies overlap.
            uint32_t am[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j)
                am[j] = __vmaxu2(__vmaxu2(cx[j].x & 0x7FFF7FFFu, cx[j].y & 0x7FFF7FFFu),
                                 __vmaxu2(cx[j].z & 0x7FFF7FFFu, cx[j].w & 0x7FFF7FFFu));
            if (p.M > CH) {
                // More than half the token capacity: every chunk holds a live token.  Branch-free so that the CH
                // independent chains (4 shuffle levels, 6 dependent conversions each) interleave; with the
                // warp-uniform guards of the path below they ran back to back (5000 cycles per 16-token group,
                // tools/tc_timeline.py — the unit cadence of the 16-token variant).  Only the stores are predicated.
#pragma unroll
                for (int lvl = 1; lvl < 16; lvl <<= 1) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) am[j] = __vmaxu2(am[j], __shfl_xor_sync(0xffffffffu, am[j], lvl));
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int m = 2 * j + tsub;
                    const uint32_t amax = min(max(am[j] & 0xFFFFu, am[j] >> 16), 0x7F7Fu);
                    const int e = (amax != 0) ? max(-100, min(100, (int)(amax >> 7) - 127 - 7)) : 0;
                    uint2 hi, lo;
                    split_e4m3x8(cx[j], __int_as_float((127 - e) << 23), hi, lo);
                    poison_nonfinite(cx[j], hi);                 // no-op for finite values
                    if (m < p.M) {
                        uint8_t* row = bstage + (m >> 3) * 1024 + (m & 7) * 128 + ((((seg8 >> 1) ^ (m & 7)) & 7) << 4) + (seg8 & 1) * 8;
                        *reinterpret_cast<uint2*>(row) = hi;
                        *reinterpret_cast<uint2*>(row + (HALF >> 3) * 1024) = lo;
                        if (seg8 == 0) xs_slot[m] = __int_as_float((127 + e) << 23);
                    }
                }
            } else {
            uint32_t nonfinite = 0;
#pragma unroll
            for (int j = 0; j < CH; ++j)
                nonfinite |= (uint32_t)(((am[j] & 0x7F80u) == 0x7F80u) | ((am[j] & 0x7F800000u) == 0x7F800000u)) << j;
#pragma unroll
            for (int lvl = 1; lvl < 16; lvl <<= 1) {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    if (2 * j < p.M) am[j] = __vmaxu2(am[j], __shfl_xor_sync(0xffffffffu, am[j], lvl));
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (2 * j < p.M) {                              // warp-uniform: chunks past the last token are skipped
                    const int m = 2 * j + tsub;
                    const uint32_t amax = min(max(am[j] & 0xFFFFu, am[j] >> 16), 0x7F7Fu);     // finite range
                    int e = 0;                                  // absmax * 2^-e in [2^7, 2^8)
                    if (amax != 0) e = max(-100, min(100, (int)(amax >> 7) - 127 - 7));
                    uint2 hi, lo;
                    split_e4m3x8(cx[j], __int_as_float((127 - e) << 23), hi, lo);
                    if ((nonfinite >> j) & 1) poison_nonfinite(cx[j], hi);
                    if (m < p.M) {
                        uint8_t* row = bstage + (m >> 3) * 1024 + (m & 7) * 128 + ((((seg8 >> 1) ^ (m & 7)) & 7) << 4) + (seg8 & 1) * 8;
                        *reinterpret_cast<uint2*>(row) = hi;
                        *reinterpret_cast<uint2*>(row + (HALF >> 3) * 1024) = lo;
                        if (seg8 == 0) xs_slot[m] = __int_as_float((127 + e) << 23);
                    }
                }
            }
            }
            fence_proxy_async_smem();                            // generic writes -> visible to the MMA's async reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(full_bar(s)); TC_PROF(4 + (g & 1)); }
#pragma unroll 1
            for (int q = 0; q < ustride && cur.valid(p); ++q) cur.next(p, G);
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> FP32 promotion -> BF16 rows / split-K fix-up =====
        const int r = tid - 128;                                 // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
        griddep_wait();
        float acc[HALF];
#pragma unroll
        for (int t = 0; t < HALF; ++t) acc[t] = 0.0f;

        // FP4 group scales: this thread's (row, group) scalars arrive through a private cp.async ring,
        // fetched kScBatch units at a time, one batch ahead (one DRAM sector serves 8 consecutive groups
        // of a row); the two scalars of a unit are adjacent in memory and come as one 8-byte copy.
        const uint32_t scslot0 = smem_u32(g_scraw) + r * 4;
        Cursor sc = cur;                                         // scale-fetch cursor, runs one batch ahead
        auto scale_fetch_batch = [&](int i0) {
            if constexpr (kIsFp4) {
#pragma unroll 1
                for (int q = 0; q < kScBatch && sc.valid(p); ++q) {
                    const int row = prow_of(sc.tile, sc.ub) + r;
                    if (r < p.R && row < p.N) {
                        const float* sp = p.scales + (size_t)row * KB + kbu_of(sc.ub) * kGroups;
#pragma unroll
                        for (int g = 0; g < kGroups; ++g)
                            if (kbu_of(sc.ub) * kGroups + g < KB)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                                             :: "r"(scslot0 + (((i0 + q) * kGroups + g) % kScDepth) * (kTileRows * 4)), "l"(sp + g)
                                             : "memory");
                    }
                    sc.next(p, G);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        };
        scale_fetch_batch(0);

        auto store_row = [&](const float (&v)[HALF], int tile_) {
            const int row = tile_ * p.R + r;
            if (r < p.R && row < p.N) {
                float rs = 1.0f, bv = 0.0f;
                if constexpr (!kIsFp4) rs = __ldg(p.scales + row);
                if (p.bias) bv = __bfloat162float(p.bias[row]);
#pragma unroll
                for (int t = 0; t < HALF; ++t)
                    if (t < p.M) p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(fmaf(v[t], rs, bv));
            }
        };

        // Row-parallel tensor parallelism: this rank holds a K slice, `v` is its FP32 partial of the tile's rows.
        // One-shot all-reduce over NVLink peer memory, fused here, with a low-latency (LL) wire format: every
        // value travels as one 8-byte store {f32 bits, epoch} into the peer's exchange buffer (slot [parity]
        // [this rank][token][row]); the receiver spins on the word itself until the epoch matches — no fence,
        // no separate flag, one NVLink one-way latency.  Each thread pushes and collects its own row, adds the P
        // partials in rank order (identical bits on every rank) and writes the BF16 row.
        // The epoch of a tile is a counter in local memory that only the tile's owner CTA advances, once per
        // call; every rank issues the same sequence of calls, so epochs agree without any host coordination
        // and the protocol replays unchanged inside a CUDA graph.  Slots alternate with the epoch's parity: a
        // rank can be at most one call ahead of a peer (it needs the peer's words of the call in between).
        auto finish_rows = [&](float (&v)[HALF], int tile_) {
            if (p.tp.world <= 1) { store_row(v, tile_); return; }
            const TpExchange& tp = p.tp;
            const int row = tile_ * p.R + r;
            const bool live = (r < p.R && row < p.N);
            // per-ROW epoch, owned by the one thread that finishes the row (the previous call's store is complete: this
            // role passed griddepcontrol.wait).  Counting per row, not per tile, keeps the tags monotone for every row
            // whatever tile height successive calls pick.
            uint32_t epoch = 0;
            if (live) { epoch = __ldcg(tp.row_epoch + row) + 1u; __stcg(tp.row_epoch + row, epoch); }
            const size_t slot_w = (size_t)kMaxTok * tp.nmax;                // words per (parity, source) slot
            float sum[HALF];
#pragma unroll
            for (int t = 0; t < HALF; ++t) sum[t] = 0.0f;
            if (live) {
                const size_t mine = ((size_t)(epoch & 1u) * tp.world + tp.rank) * slot_w + row;
                for (int q = 0; q < tp.world; ++q) {
                    if (q == tp.rank) continue;
                    uint2* dst = tp.data[q] + mine;
#pragma unroll
                    for (int t = 0; t < HALF; ++t)
                        if (t < p.M)
                            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};"
                                         :: "l"(dst + (size_t)t * tp.nmax), "r"(__float_as_uint(v[t])), "r"(epoch) : "memory");
                }
                const long long t0 = clock64();
                for (int q = 0; q < tp.world; ++q) {
                    const uint2* src = tp.data[tp.rank] + ((size_t)(epoch & 1u) * tp.world + q) * slot_w + row;
#pragma unroll
                    for (int t = 0; t < HALF; ++t) {
                        if (t < p.M) {
                            if (q == tp.rank) { sum[t] += v[t]; continue; }
                            uint32_t bits, tag;
                            do {
                                asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];"
                                             : "=r"(bits), "=r"(tag) : "l"(src + (size_t)t * tp.nmax) : "memory");
                                if (tag != epoch && clock64() - t0 > 40000000000LL) __trap();   // ~20 s: a peer died
                            } while (tag != epoch);
                            sum[t] += __uint_as_float(bits);
                        }
                    }
                }
            }
            store_row(sum, tile_);
        };

        float gate[HALF];                                       // fused GLU: BF16-rounded gate projections of the tile
#pragma unroll
        for (int t = 0; t < HALF; ++t) gate[t] = 0.0f;
        const int first_tile = cur.tile;                         // stream-K: tile of this CTA's first segment
        int seg_first_ub = cur.ub;
        for (int i = 0; cur.valid(p); ++i) {
            const int slot = i % kTmemUnits, tph = (i / kTmemUnits) & 1;
            if constexpr (kIsFp4) {
                if ((i % kScBatch) == 0) {                       // batch i/kScBatch is needed now: fetch the next one
                    scale_fetch_batch(i + kScBatch);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                }
            }
            if (r == 0) TC_PROF(9);
            mbar_wait(tfull_bar(slot), tph);
            if (r == 0) TC_PROF(10);
            tcgen05_fence_after();
            uint32_t d[kGroups][NCOLS];
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const uint32_t ta = tmem_base + lane_base + (slot * kGroups + g) * NCOLS;
                if constexpr (NCOLS == 16) tmem_ld_32x32b_x16(ta, d[g]);
                else                       tmem_ld_32x32b_x32(ta, d[g]);
            }
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(slot));
            if (r == 0) TC_PROF(11);

#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                float wsc = 1.0f;
                if constexpr (kIsFp4) wsc = g_scraw[((i * kGroups + g) % kScDepth) * kTileRows + r];
                const float4* xs4 = reinterpret_cast<const float4*>(g_xs + ((i * kGroups + g) % kXsRing) * kMaxTok);
#pragma unroll
                for (int q = 0; q < HALF / 4; ++q) {
                    const float4 xs = xs4[q];
                    const float xv[4] = { xs.x, xs.y, xs.z, xs.w };
                    if constexpr (NCOLS == 32) {
                        // 9..16 tokens: this loop is the kernel's cadence (profiles/r2k6_*): packed FP32 pairs (FFMA2) — the same
                        // operations and roundings, half the instructions
#pragma unroll
                        for (int j = 0; j < 4; j += 2) {
                            const int t = q * 4 + j;
                            float dv0, dv1;
                            fma_f32x2(dv0, dv1, __uint_as_float(d[g][HALF + t]), __uint_as_float(d[g][HALF + t + 1]), 0.0625f, 0.0625f,
                                      __uint_as_float(d[g][t]), __uint_as_float(d[g][t + 1]));
                            if constexpr (kIsFp4) {
                                float p0, p1;
                                mul_f32x2(p0, p1, dv0, dv1, xv[j], xv[j + 1]);
                                fma_f32x2(acc[t], acc[t + 1], p0, p1, wsc, wsc, acc[t], acc[t + 1]);
                            } else {
                                fma_f32x2(acc[t], acc[t + 1], dv0, dv1, xv[j], xv[j + 1], acc[t], acc[t + 1]);
                            }
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int t = q * 4 + j;
                        const float dv = fmaf(__uint_as_float(d[g][HALF + t]), 0.0625f, __uint_as_float(d[g][t]));
                        if constexpr (kIsFp4) acc[t] = fmaf(dv * xv[j], wsc, acc[t]);
                        else                  acc[t] = fmaf(dv, xv[j], acc[t]);
                    }
                    }
                }
            }

            if (p.glu && cur.ub == KBH - 1) {
                // end of the gate rows: round the gate projection to BF16 exactly as the unfused Linear stores it
                const int grow = min(cur.tile * p.R + r, p.H - 1);       // (rows past the tile / past H are never stored)
                const float rs = (kIsFp4 ? 1.0f : __ldg(p.scales + grow));
                const float bv = p.bias ? __bfloat162float(p.bias[grow]) : 0.0f;
#pragma unroll
                for (int t = 0; t < HALF; ++t) { gate[t] = bf16_round(fmaf(acc[t], rs, bv)); acc[t] = 0.0f; }
            } else if (p.glu && cur.item_end(p)) {
                const int hrow = cur.tile * p.R + r, urow = p.H + min(hrow, p.H - 1);
                const bool live = (r < p.R && hrow < p.H);
                const float rs = (kIsFp4 ? 1.0f : __ldg(p.scales + urow));
                const float bv = p.bias ? __bfloat162float(p.bias[urow]) : 0.0f;
                if (p.M > 2) {
                    // several tokens: all HALF activations straight-line (glu.cuh), then the stores
                    float up[HALF];
                    __nv_bfloat16 o[HALF];
#pragma unroll
                    for (int t = 0; t < HALF; ++t) { up[t] = bf16_round(fmaf(acc[t], rs, bv)); acc[t] = 0.0f; }
                    glu_combine_many(p.glu, gate, up, o);
#pragma unroll
                    for (int t = 0; t < HALF; ++t)
                        if (live && t < p.M) p.y[(size_t)t * p.H + hrow] = o[t];
                } else {
#pragma unroll
                    for (int t = 0; t < HALF; ++t) {
                        if (live && t < p.M) p.y[(size_t)t * p.H + hrow] = glu_combine(p.glu, gate[t], bf16_round(fmaf(acc[t], rs, bv)));
                        acc[t] = 0.0f;
                    }
                }
            } else
            if (cur.item_end(p)) {
                const int tile = cur.tile;
                // is the finished segment the whole tile, and if not, which workspace slot is ours?
                const bool whole = cur.sk ? (seg_first_ub == 0 && cur.ub == p.KBU - 1) : (p.P == 1);
                if (whole) {
                    finish_rows(acc, tile);                                // every k of these rows was ours
                } else {
                    // split fix-up: park the partial, take a ticket; the last contributor adds all partials in a
                    // fixed order (item j / CTA index) and writes the rows — same bits every run
                    if (p.cl) {
                        // Cluster fix-up: the P CTAs of this tile are one thread-block cluster (rank = k-split index).
                        // Every CTA's weight stages are idle by now (its last MMA has been consumed), so a non-leader
                        // parks its partial in its OWN stage 0, releases it with a cluster-scope arrive on the leader's
                        // mbarrier and stays alive until the leader has pulled it over distributed shared memory; the
                        // leader adds the partials in rank order.  One DSMEM round trip instead of a global-memory
                        // fence + atomic ticket + L2 reads (~2 us at the tail of every split launch).
                        const uint32_t rank = cluster_ctarank();
                        float* xbuf = reinterpret_cast<float*>(smem_raw);
                        const uint32_t xbar = smem_u32(g_misc + 16), dbar = smem_u32(g_misc + 24);
                        if (rank != 0) {
#pragma unroll
                            for (int t = 0; t < HALF; ++t)
                                if (t < p.M) xbuf[t * kTileRows + r] = acc[t];
                            mbar_arrive_release_cluster(mapa_shared(xbar, 0));
                            if (r == 0) mbar_wait(dbar, 0);                 // the leader has read our partial
                            bar_sync(1, 128);
                        } else {
                            mbar_wait_acquire_cluster(xbar, 0);
                            float v[HALF];
#pragma unroll
                            for (int t = 0; t < HALF; ++t) v[t] = acc[t];
                            for (int j = 1; j < p.P; ++j) {
                                const uint32_t src = mapa_shared(smem_u32(xbuf) + r * 4, j);
#pragma unroll
                                for (int t = 0; t < HALF; ++t)
                                    if (t < p.M) v[t] += ld_shared_cluster_f32(src + t * kTileRows * 4);
                            }
                            bar_sync(1, 128);                               // every row has been pulled
                            if (r >= 1 && r < p.P) mbar_arrive_cluster(mapa_shared(dbar, r));
                            finish_rows(v, tile);
                        }
                    } else {
                    const int my_slot = cur.sk ? 2 * (int)blockIdx.x + (tile != first_tile ? 1 : 0) : cur.it;
                    float* wp = p.ws + (size_t)my_slot * kWsSlotFloats + r;
#pragma unroll
                    for (int t = 0; t < HALF; ++t)
                        if (t < p.M) __stcg(wp + t * kTileRows, acc[t]);
                    __threadfence();
                    bar_sync(1, 128);
                    const long long T = (long long)p.tiles * p.KBU;
                    if (r == 0) {
                        int need = p.P;
                        if (cur.sk) need = sk_owner((tile + 1) * p.KBU - 1, T, G) - sk_owner(tile * p.KBU, T, G) + 1;
                        const int old = atomicAdd(p.counters + tile, 1);
                        *g_flag = (old == need - 1);
                    }
                    bar_sync(1, 128);
                    const bool last = (*g_flag != 0);
                    bar_sync(1, 128);                                       // flag consumed before any rewrite
                    if (last) {
                        __threadfence();
                        float v[HALF];
#pragma unroll
                        for (int t = 0; t < HALF; ++t) v[t] = 0.0f;
                        const int j0 = cur.sk ? sk_owner(tile * p.KBU, T, G) : 0;
                        const int j1 = cur.sk ? sk_owner((tile + 1) * p.KBU - 1, T, G) : p.P - 1;
                        for (int j = j0; j <= j1; ++j) {
                            const int slot = cur.sk ? 2 * j + (tile != sk_start(j, T, G) / p.KBU ? 1 : 0) : tile * p.P + j;
                            const float* rp = p.ws + (size_t)slot * kWsSlotFloats + r;
#pragma unroll
                            for (int t = 0; t < HALF; ++t)
                                if (t < p.M) v[t] += __ldcg(rp + t * kTileRows);
                        }
                        finish_rows(v, tile);
                        if (r == 0) p.counters[tile] = 0;                   // ready for the next launch
                    }
                    }
                }
#pragma unroll
                for (int t = 0; t < HALF; ++t) acc[t] = 0.0f;
                seg_first_ub = (cur.ub == p.KBU - 1) ? 0 : cur.ub + 1;
            }
            cur.next(p, G);
        }
        if (r == 0) TC_PROF_CTA(2);
    }

    // ---- teardown ----------------------------------------------------------------------------------
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (tid == 0) {
        TC_PROF_CTA(3);
        if constexpr (PROF) {
            if (p.prof) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.prof[1024 + 148 * 4 + blockIdx.x] = sm; }
        }
    }
}

// ---- activation pre-pass (M > 8) -------------------------------------------------------------------
// With more than 8 tokens the in-kernel converter warps set the unit cadence (every CTA re-converts the same
// M x K activations for its own row tile: 5000 warp-cycles per 16-token group against a 700-cycle unit), so the
// split is done ONCE per forward by this kernel and the decode kernel's producer bulk-copies the result.
// One CTA per 128-k group, warp j = tokens 2j and 2j + 1, lane = (token parity, 8-k segment): the arithmetic
// and the bytes are exactly those of the converter warps above.  Output: for every group the NCOLS x 128-byte
// shared-memory image of the B operand (hi rows, lo rows, 128B-swizzled — a 1-D bulk copy drops it into a stage)
// and kMaxTok block scales.  Token rows >= M and groups >= KB (padding of the last unit) are written as zeros.
template <int NCOLS>
__global__ void __launch_bounds__(512)
act_presplit_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ img, float* __restrict__ xs,
                    int M, int K, int KB, const NormArgs norm)
{
    constexpr int HALF = NCOLS / 2;
    griddep_launch_dependents();                // the decode kernel may start streaming its weights
    griddep_wait();                             // x is the previous kernel's output; it may also still read img
    const int kb = blockIdx.x, j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg8 = lane & 15, tsub = lane >> 4, m = 2 * j + tsub;
    const bool live = (m < M && kb < KB);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (live) v = __ldcg(reinterpret_cast<const uint4*>(x + (size_t)m * K + (size_t)kb * kBlockK + seg8 * 8));
    if (norm.on) {
        // fused RMSNorm: the launch then has 16 warps, warp w computes the reciprocal RMS of token w in the reference's
        // reduction order (one sequential FMA chain per lane: ~1 us — two tokens per warp doubled that on the dependency
        // chain of every Linear); warps 8-15 are done after that
        __shared__ float s_rstd[kMaxTok];
        if (j < M && j < kMaxTok) {
            const float rs = rms_rstd_select(norm, x + (size_t)j * K, K, lane);
            if (lane == 0) s_rstd[j] = rs;
        }
        __syncthreads();
        if (threadIdx.x >= NCOLS * 8) return;
        const float rs0 = (2 * j < M) ? s_rstd[2 * j] : 1.0f, rs1 = (2 * j + 1 < M) ? s_rstd[2 * j + 1] : 1.0f;
        if (live) {
            uint4 w8 = make_uint4(0, 0, 0, 0), b8 = make_uint4(0, 0, 0, 0);
            if (norm.weight) w8 = __ldg(reinterpret_cast<const uint4*>(norm.weight + (size_t)kb * kBlockK + seg8 * 8));
            if (norm.bias) b8 = __ldg(reinterpret_cast<const uint4*>(norm.bias + (size_t)kb * kBlockK + seg8 * 8));
            v = rms_apply8(v, tsub ? rs1 : rs0, w8, b8, norm.weight != nullptr, norm.bias != nullptr, norm.weight_offset);
        }
    }
    uint32_t am = __vmaxu2(__vmaxu2(v.x & 0x7FFF7FFFu, v.y & 0x7FFF7FFFu), __vmaxu2(v.z & 0x7FFF7FFFu, v.w & 0x7FFF7FFFu));
#pragma unroll
    for (int lvl = 1; lvl < 16; lvl <<= 1) am = __vmaxu2(am, __shfl_xor_sync(0xffffffffu, am, lvl));
    const uint32_t amax = min(max(am & 0xFFFFu, am >> 16), 0x7F7Fu);
    const int e = (amax != 0) ? max(-100, min(100, (int)(amax >> 7) - 127 - 7)) : 0;
    uint2 hi, lo;
    split_e4m3x8(v, __int_as_float((127 - e) << 23), hi, lo);
    poison_nonfinite(v, hi);                    // no-op for finite values
    uint8_t* row = img + (size_t)kb * (NCOLS * 128) + (m >> 3) * 1024 + (m & 7) * 128 + ((((seg8 >> 1) ^ (m & 7)) & 7) << 4) + (seg8 & 1) * 8;
    *reinterpret_cast<uint2*>(row) = hi;        // zeros for a dead row: split(0) = (0, 0)
    *reinterpret_cast<uint2*>(row + (HALF >> 3) * 1024) = lo;
    if (seg8 == 0 && m < kMaxTok) xs[(size_t)kb * kMaxTok + m] = live ? __int_as_float((127 + e) << 23) : 0.0f;
}

// =================================================================================================
// host side
// =================================================================================================

struct MapKey {
    const void* ptr; int N, K, fmt, R;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && N == o.N && K == o.K && fmt == o.fmt && R == o.R; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= ((size_t)k.N << 32) ^ ((size_t)k.K << 10) ^ ((size_t)k.R << 2) ^ (size_t)k.fmt;
        return h;
    }
};

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

// The tensor map of a weight matrix: [N rows, K elements], box = 128 k x R rows, 128B swizzle.  (Also used by decode_chain.cu.)
int weight_tensor_map(const void* w, int N, int K, int fmt, int R, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ w, N, K, fmt, R };
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return MILAB200_E_NO_DEVICE;
    const bool fp4 = (fmt != kFp8);
    const cuuint64_t dims[2] = { (cuuint64_t)K, (cuuint64_t)N };
    const cuuint64_t strides[1] = { (cuuint64_t)(fp4 ? K / 2 : K) };
    const cuuint32_t box[2] = { (cuuint32_t)kBlockK, (cuuint32_t)R };
    const cuuint32_t estr[2] = { 1, 1 };
    const int promo = env_int("MILAB200_TMA_L2_PROMO", 3);
    const CUresult r = enc(out, fp4 ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                           const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MILAB200_E_BAD_SHAPE;
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

namespace {

struct TcDevice {
    bool ready = false, failed = false;
    int sms = 0;
    float* ws = nullptr;          // kWsRegions x (sms*2) partial-tile slots
    int* counters = nullptr;      // kWsRegions x kMaxTiles
    uint8_t* ps_img = nullptr;    // kWsRegions x kPsMaxGroups x 4 KB: pre-split activation planes (M > 8)
    float* ps_xs = nullptr;       // kWsRegions x kPsMaxGroups x kMaxTok block scales
    std::atomic<unsigned> next_region{0};
    int max_cl[9] = {};           // co-resident thread-block clusters of P decode CTAs (1 CTA per SM), P = 2..8
};
constexpr int kPsMaxGroups = 1024;       // K <= 131072 through the pre-split path
constexpr int kPsImgBytes = 32 * 128;    // one group of the 16-token variant
TcDevice g_tc[16];
std::mutex g_tc_mu;
std::atomic<bool> g_tc_enabled{ env_int("MILAB200_DECODE_TC", 1) != 0 };
std::atomic<bool> g_weights_fresh[16];
std::atomic<int> g_streamk{ env_int("MILAB200_STREAMK", -1) };
std::atomic<int> g_presplit{ env_int("MILAB200_PRESPLIT", 1) };      // M > 8: activation pre-pass (1) or converter warps (0)
long long* g_tc_prof = nullptr;          // bring-up timeline buffer (tools/tc_timeline.py), normally null   // a kernel of this library wrote weight storage since the last decode launch

// Allocates the stream-K workspace of the current device on first use.  Allocation is not legal
// while a stream is being captured, so callers capture only after one eager call (or
// milab200_init()); inside a capture an unprepared device reports "not ready".
TcDevice* tc_device(cudaStream_t stream)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    TcDevice& d = g_tc[dev];
    if (d.ready) return &d;
    if (d.failed) return nullptr;
    std::lock_guard<std::mutex> lk(g_tc_mu);
    if (d.ready) return &d;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cs != cudaStreamCaptureStatusNone) return nullptr;
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10 || d.sms <= 0 || !encode_tiled_fn()) { d.failed = true; return nullptr; }
    const size_t ws_bytes = (size_t)kWsRegions * kMaxSplitItems * kWsSlotFloats * sizeof(float);
    const size_t ct_bytes = (size_t)kWsRegions * kMaxTiles * sizeof(int);
    const size_t pi_bytes = (size_t)kWsRegions * kPsMaxGroups * kPsImgBytes;
    const size_t px_bytes = (size_t)kWsRegions * kPsMaxGroups * kMaxTok * sizeof(float);
    if (cudaMalloc(&d.ws, ws_bytes) != cudaSuccess || cudaMalloc(&d.counters, ct_bytes) != cudaSuccess ||
        cudaMalloc(&d.ps_img, pi_bytes) != cudaSuccess || cudaMalloc(&d.ps_xs, px_bytes) != cudaSuccess ||
        cudaMemset(d.counters, 0, ct_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        d.failed = true;
        return nullptr;
    }
    // how many clusters of P decode CTAs can be co-resident (GPCs of 16/18/20 SMs strand some SMs for P > 2)
    if (cudaFuncSetAttribute(decode_tc_kernel<kFp8, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)TcShape<16>::kSmem) == cudaSuccess) {
        for (int P = 2; P <= 8; ++P) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(P * (d.sms / P)); cfg.blockDim = dim3(TcShape<16>::kThreads); cfg.dynamicSmemBytes = TcShape<16>::kSmem;
            cudaLaunchAttribute a[1];
            a[0].id = cudaLaunchAttributeClusterDimension;
            a[0].val.clusterDim.x = P; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
            cfg.attrs = a; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, decode_tc_kernel<kFp8, 16, false>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
            d.max_cl[P] = n;
        }
    } else cudaGetLastError();
    d.ready = true;
    return &d;
}

// k-splits per row tile.  Whole row tiles (P = 1) need no cross-CTA reduction and are preferred as soon
// as there are enough of them to keep ~3/4 of the SMs streaming (the kernel is HBM-bound: ~110 SMs
// saturate the memory system).  With few row tiles the k range is cut so that one wave of items covers
// most SMs; each item stays >= 4 k blocks.
int choose_split(int tiles, int KB, int sms)
{
    static const int forced = env_int("MILAB200_SPLITK", 0);
    int P = 1;
    if (forced > 0) P = forced;
    else if (tiles * 4 < sms * 3) {
        P = sms / tiles;                                   // one wave: tiles * P <= sms
        while (P > 1 && KB / P < 2) --P;                   // (KB counts 256-k unit blocks here)
    }
    if (P > KB) P = KB;
    if (P < 1) P = 1;
    while (P > 1 && tiles * P > kMaxSplitItems) --P;
    return P;
}

// Balanced decomposition: tile height R (<= 128 rows) and k-splits P chosen together so that ONE wave of
// ceil(rows / R) * P items covers (nearly) every SM with equal bytes.  The decode kernels stream ~47 GB/s per SM
// whatever the shape (profiles/README.md), so a launch lasts as long as its busiest SM: 112 tiles of 128 rows
// (Llama-8B gate: 14336 rows) keep 112 of 148 SMs busy for 8 units each, 148 tiles of 97 rows keep all of them busy
// for 6.06 units' worth of bytes — with no cross-CTA reduction at all, because a row tile still owns its whole k
// range.  The MMA is always 128 rows tall; a shorter box just leaves the upper TMEM lanes unused.
//   cost(R, P) = waves * units-per-item * (R + c_unit) [+ c_fix when P > 1]   in units of one row x 512 k (~10 ns)
// P > 1 only as a single wave of co-resident clusters (DSMEM fix-up); max_cl[P] = co-resident clusters of P CTAs.
// Measured (profiles/r2j1_ab_tile_rows.txt): for a stand-alone launch the balanced wave only pays when whole 128-row
// tiles need more than one wave (Gemma gate_up at M = 16: 26.3 -> 23.5 us).  With <= one wave of tiles a launch is
// bounded by its fixed ~2.8 us of ramp / dependency bubble, not by bytes per SM, and idle SMs are where the NEXT
// kernel's CTAs prefetch their weights under programmatic dependent launch (bench: 876 vs 819 tok/s) — so those
// shapes keep whole tiles (`balanced` = false).  The chained kernel (one launch walking many Linears) has no such
// bubble and always balances.
struct Decomp { int R, P; };
Decomp choose_decomp(int rows, int KBU, int sms, bool allow_split, const int* max_cl, bool balanced)
{
    static const int forced_r = env_int("MILAB200_TILE_ROWS", 0);       // 0 = auto; 128 = whole 128-row tiles; -1 = always balanced
    if (forced_r == 0 && !balanced) {
        const int tiles = (rows + kTileRows - 1) / kTileRows;
        return { kTileRows, allow_split ? choose_split(tiles, KBU, sms) : 1 };
    }
    static const int forced_p = env_int("MILAB200_SPLITK", 0);
    static const int c_unit = env_int("MILAB200_COST_UNIT", 8), c_fix = env_int("MILAB200_COST_FIXUP", 48);
    if (forced_r > 0) {
        const int R = forced_r > kTileRows ? kTileRows : forced_r;
        const int tiles = (rows + R - 1) / R;
        return { R, allow_split ? choose_split(tiles, KBU, sms) : 1 };
    }
    // memo: the search below is ~1000 integer steps, the answer depends on (rows, KBU, allow_split) only
    static std::mutex mu;
    static std::unordered_map<uint64_t, Decomp> memo;
    const uint64_t key = ((uint64_t)(uint32_t)rows << 32) | ((uint64_t)(uint32_t)KBU << 1) | (allow_split ? 1u : 0u);      // (balanced only)
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = memo.find(key);
        if (it != memo.end()) return it->second;
    }
    Decomp best{ kTileRows, 1 };
    long long best_cost = -1;
    for (int P = 1; P <= (allow_split ? 8 : 1); ++P) {
        if (forced_p > 0 && P != forced_p) continue;
        if (P > 1 && (KBU / P < 2 || max_cl[P] <= 0)) continue;
        const int upi = (KBU + P - 1) / P;
        for (int R = 16; R <= kTileRows; ++R) {
            const int tiles = (rows + R - 1) / R;
            if (tiles > kMaxTiles) continue;
            const long long items = (long long)tiles * P;
            if (P > 1 && (items > sms || tiles > max_cl[P] || items > kMaxSplitItems)) continue;
            const long long waves = (items + sms - 1) / sms;
            const long long cost = waves * upi * (R + c_unit) + (P > 1 ? c_fix : 0);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = { R, P }; }
        }
    }
    std::lock_guard<std::mutex> lk(mu);
    if (memo.size() > 4096) memo.clear();
    memo.emplace(key, best);
    return best;
}

static_assert(TcShape<16>::kSmem <= 232448 && TcShape<32>::kSmem <= 232448, "exceeds 227 KB of shared memory per CTA");

template <int FMT, int NCOLS, bool PROF>
int launch_tc_impl(const CUtensorMap& tm, const TcParams& p, int grid, cudaStream_t stream, const char* name, bool allow_pdl)
{
    constexpr size_t smem = TcShape<NCOLS>::kSmem;
    static std::atomic<bool> configured[16];
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !configured[dev].load()) {
        MILAB200_RETURN_IF_CUDA(cudaFuncSetAttribute(decode_tc_kernel<FMT, NCOLS, PROF>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev].store(true);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TcShape<NCOLS>::kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int nattr = 0;
    static const int pdl = env_int("MILAB200_PDL", 1);
    if (pdl && allow_pdl) {
        attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[nattr].val.programmaticStreamSerializationAllowed = 1;
        ++nattr;
    }
    TcParams q = p;
    if (p.cl) {
        // the P k-splits of a tile as one thread-block cluster: only if that many clusters can be co-resident
        static std::atomic<int> max_clusters[16][9];            // per device, per cluster size (0 = not queried yet)
        attr[nattr].id = cudaLaunchAttributeClusterDimension;
        attr[nattr].val.clusterDim.x = p.P; attr[nattr].val.clusterDim.y = 1; attr[nattr].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = nattr + 1;
        int mc = (dev >= 0 && dev < 16) ? max_clusters[dev][p.P].load() : 0;
        if (mc == 0) {
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, decode_tc_kernel<FMT, NCOLS, PROF>, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
            mc = n > 0 ? n : -1;
            if (dev >= 0 && dev < 16) max_clusters[dev][p.P].store(mc);
        }
        if (mc >= p.tiles) ++nattr; else q.cl = 0;               // else: the global-memory fix-up
    }
    cfg.attrs = attr; cfg.numAttrs = nattr;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, decode_tc_kernel<FMT, NCOLS, PROF>, tm, q);
    if (e != cudaSuccess) return (int)e;
    note_launch(name);
    return 0;
}

template <int FMT, int NCOLS>
int launch_tc(const CUtensorMap& tm, const TcParams& p, int grid, cudaStream_t stream, const char* name, bool allow_pdl)
{
#ifdef MILAB200_DIAG
    if (p.prof) return launch_tc_impl<FMT, NCOLS, true>(tm, p, grid, stream, name, allow_pdl);    // role-timeline build (make diag)
#endif
    return launch_tc_impl<FMT, NCOLS, false>(tm, p, grid, stream, name, allow_pdl);
}

}  // namespace

// Returns 1 when the shape / device is not eligible (the caller takes the mma.sync kernels), else 0
// with the launch status in *status.
int try_decode_tc_norm(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                       const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status,
                       const TpExchange* tp, int glu, const NormArgs* norm);
int try_decode_tc(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                  const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status,
                  const TpExchange* tp, int glu)
{
    return try_decode_tc_norm(fmt, y, x, w, scales, bias, M, K, N, stream, status, tp, glu, nullptr);
}

int try_decode_tc_norm(int fmt, __nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                       const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status,
                       const TpExchange* tp, int glu, const NormArgs* norm)
{
    if (!g_tc_enabled.load(std::memory_order_relaxed)) return 1;
    if (fmt != kFp8 && fmt != kFp4G128) return 1;
    if (M < 1 || M > kMaxTok || K % kBlockK != 0) return 1;
    const uintptr_t wa = reinterpret_cast<uintptr_t>(w);
    if ((wa & 31) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
    // fused GLU: N = 2 H physical rows (gate | up), one logical tile = 128 gate rows + the 128 up rows H below
    if (glu && (tp || N % 2 != 0)) return 1;
    TcDevice* d = tc_device(stream);
    if (!d) return 1;
    const int groups = (M <= 8) ? TcShape<16>::kGroups : TcShape<32>::kGroups;
    const int KBU1 = (K / kBlockK + groups - 1) / groups;        // units of one row tile (of one GLU half)
    const int rows_l = glu ? N / 2 : N;
    const bool multi_wave = (rows_l + kTileRows - 1) / kTileRows > d->sms;
    const Decomp dc = choose_decomp(rows_l, glu ? 2 * KBU1 : KBU1, d->sms, !glu, d->max_cl,
                                    multi_wave && (M <= 8 || KBU1 < 24));
    const int R = dc.R;
    const int tiles = ((glu ? N / 2 : N) + R - 1) / R;
    if (tiles > kMaxTiles) return 1;
    if (glu && tiles * 4 < d->sms * 3) return 1;
    // M > 8 on a badly unbalanced two-wave shape: the unfused Linear (stream-K) + activation kernel is faster (measured)
    if (glu && M > 8 && tiles > d->sms && (long long)((tiles + d->sms - 1) / d->sms) * d->sms * 5 >= (long long)tiles * 6) return 1;          // whole logical tiles only (no split): needs enough of them

    CUtensorMap tm;
    const int rc = weight_tensor_map(w, N, K, fmt, R, &tm);
    if (rc != 0) return 1;

    TcParams p;
    p.y = y; p.x = x; p.scales = scales; p.bias = bias;
    p.norm = norm ? *norm : NormArgs();
    p.M = M; p.K = K; p.N = N; p.KB = K / kBlockK; p.KBU = KBU1; p.tiles = tiles; p.R = R;
    // Stream-K (P = 0) balances every SM to the same number of units, but each cut tile pays a ~2 us fix-up
    // (__threadfence + ticket + partial reads) at the kernel's tail; measured on one box it loses to whole-tile
    // items at M <= 8 on every shape (Llama-8B gate 13.8 vs 11.4 us) and wins only where whole tiles leave a
    // badly unbalanced second wave AND the kernel is not HBM-bound (Llama-70B shapes at M = 16: 55.9 vs 73.0 us).
    const int streamk_mode = g_streamk.load(std::memory_order_relaxed);      // -1 auto, 0 off, 1 on
    const int waves = (tiles + d->sms - 1) / d->sms;
    const bool unbalanced = tiles > d->sms && (long long)waves * d->sms * 5 >= (long long)tiles * 6;
    const bool streamk = !glu && (streamk_mode == 1 || (streamk_mode < 0 && M > 8 && unbalanced));
    p.glu = glu; p.H = N / 2;
    if (glu) p.KBU *= 2;                                  // the units of the gate rows, then those of the up rows
    p.P = streamk ? 0 : dc.P;
    p.items = streamk ? tiles * p.KBU : tiles * p.P;
    static const int cluster_fixup = env_int("MILAB200_CLUSTER_FIXUP", 1);
    p.cl = (cluster_fixup && !streamk && p.P > 1 && p.P <= 8 && p.items <= d->sms) ? 1 : 0;
    const unsigned region = d->next_region.fetch_add(1) % kWsRegions;
    p.ws = d->ws + (size_t)region * kMaxSplitItems * kWsSlotFloats;
    p.counters = d->counters + (size_t)region * kMaxTiles;
    p.a_tx_bytes = (fmt == kFp8) ? (uint32_t)(R * kBlockK) : (uint32_t)(R * kBlockK / 2);   // bytes of one box as they lie in HBM
#ifdef MILAB200_DIAG
    p.prof = g_tc_prof;
#else
    p.prof = nullptr;
#endif
    static const int early_ld = env_int("MILAB200_EARLY_LD", 1);
    p.early_ld = early_ld;
    if (tp) p.tp = *tp; else p.tp.world = 1;
    if (p.tp.world > 1 && (N > p.tp.nmax || tiles > kMaxTiles)) return 1;
    const int grid = p.items < d->sms ? p.items : d->sms;
    // The TMA producer reads the weights before griddepcontrol.wait.  That is only legal when the
    // weights were complete before the previous kernel in the stream began; a launch that directly
    // follows one of this library's quantizers therefore takes ordinary stream order.
    int dev = 0; cudaGetDevice(&dev);
    const bool pdl_ok = !(dev >= 0 && dev < 16 && g_weights_fresh[dev].exchange(false));

    // M > 8: split the activations once, ahead of the decode kernel (MILAB200_PRESPLIT=0: converter warps)
    const int presplit = g_presplit.load(std::memory_order_relaxed);
    p.ps = 0; p.xp = nullptr; p.xps = nullptr;
    bool decode_pdl = pdl_ok;
    static const int prof_ps = env_int("MILAB200_PROF_PRESPLIT", 0);      // diag builds: keep the pre-pass under the role timeline
    if (presplit && M > 8 && (!p.prof || prof_ps) && p.KBU * groups <= kPsMaxGroups) {
        uint8_t* img = d->ps_img + (size_t)region * kPsMaxGroups * kPsImgBytes;
        float* pxs = d->ps_xs + (size_t)region * kPsMaxGroups * kMaxTok;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((glu ? p.KBU / 2 : p.KBU) * groups); cfg.blockDim = dim3(p.norm.on ? 512 : 32 * 8); cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        static const int pdl = env_int("MILAB200_PDL", 1);
        if (pdl && pdl_ok) {
            // (a launch right behind a quantizer keeps stream order: the decode kernel that follows prefetches
            // weights as soon as THIS kernel starts, and this kernel must then start after the quantizer has ended)
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
        }
        const cudaError_t e = cudaLaunchKernelEx(&cfg, act_presplit_kernel<32>, x, img, pxs, M, K, p.KB, p.norm);
        if (e != cudaSuccess) { *status = (int)e; return 0; }
        note_launch("act_presplit_kernel<n32>");
        p.ps = 1; p.xp = img; p.xps = pxs;
        decode_pdl = true;
    }

    if (fmt == kFp8)
        *status = (M <= 8) ? launch_tc<kFp8, 16>(tm, p, grid, stream, "decode_tc_kernel<fp8,n16>", decode_pdl)
                           : launch_tc<kFp8, 32>(tm, p, grid, stream, "decode_tc_kernel<fp8,n32>", decode_pdl);
    else
        *status = (M <= 8) ? launch_tc<kFp4G128, 16>(tm, p, grid, stream, "decode_tc_kernel<fp4g128,n16>", decode_pdl)
                           : launch_tc<kFp4G128, 32>(tm, p, grid, stream, "decode_tc_kernel<fp4g128,n32>", decode_pdl);
    return 0;
}

void tc_set_enabled(bool on) { g_tc_enabled.store(on); }
void tc_set_streamk(int mode) { g_streamk.store(mode); }
void tc_set_presplit(int on) { g_presplit.store(on); }
int tc_streamk_mode() { return g_streamk.load(); }

bool tc_take_weights_fresh()
{
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 16 && g_weights_fresh[dev].exchange(false);
}
void tc_set_prof(long long* buf) { g_tc_prof = buf; }
long long* tc_prof_buffer() { return g_tc_prof; }

void tc_note_weights_written()
{
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 16) g_weights_fresh[dev].store(true);
}

int tc_prepare_device()
{
    return tc_device(nullptr) ? 0 : MILAB200_E_NO_DEVICE;
}

}  // namespace milab200
