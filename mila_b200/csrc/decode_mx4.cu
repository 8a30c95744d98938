// decode_mx4.cu — FP4 (PerGroupFp4<128>) decode for M <= 8 tokens with the weights kept PACKED end to end:
// TMA -> shared memory -> tcgen05.mma kind::mxf4 -> TMEM -> FP32 promotion.
//
// Replaces cuda_matvec_decode_bf16_qfp4 (LIN/Kernels/MatVec/CudaMatVecBias.Bf16.cu:271,:376,:545) at M = 1 and
// serves M = 2..4 of cuda_fp4a16_gemm (K/W4A16Gemm/CudaW4A16Gemm.cu:88).
//
// Why a second FP4 kernel: decode_tc.cu feeds E2M1 to kind::f8f6f4, which wants every nibble in a byte
// container — the 16U4_ALIGN16B tensor map unpacks on the way in, so a byte of HBM costs two bytes of shared
// memory and the stages that fit (3 x 72 KB) keep only 96 KB of HBM bytes in flight per SM.  Measured
// consequence: both formats retire about one 128-row x 512-k unit per 1.3 us per SM, which is the HBM peak for
// FP8 (64 KB units) but 58 % of it for FP4 (32 KB units) — profiles/r1_perf_shapes_*.jsonl.
// kind::mxf4 is the one tcgen05 kind that reads packed nibbles (two per byte, low nibble = even k: Mila's
// layout, Policies.ixx:104-113).  It is a block-scaled kind with BOTH operands E2M1:
//   * scale factors: Mila's FP32 group scales are not UE8M0, so the hardware scaling is
//     neutralised — the scale-factor region of TMEM is filled once with 0x7F (UE8M0 2^0) for A and B — and
//     the real (row, group) scale is applied in the FP32 promotion, exactly as in decode_tc.cu.
//   * activations: each BF16 value is written as a 16-bit fixed-point magnitude relative to its token's
//     128-k block maximum, u = rn(|x| 2^(15-E)), and cut into eight 2-bit digits; digit d in {0,1,2,3} IS an
//     E2M1 number, so plane p (nibble = sign | code(d_p)) is an exact E2M1 operand and
//     x 2^(15-E) = sum_p 4^p plane_p.  Eight planes per token are eight MMA columns (N = 32 for up to 4
//     tokens); the epilogue recombines them with seven immediate-form FMAs (Horner in base 4).  Products and
//     the 128-k sums are exact in the tensor core; values below 2^-8 of their block maximum are rounded to
//     2^-16 of it (absolute error < 2^-17 of the block maximum; decode_tc.cu's floor is 2^-21).
//   * stage = 128 rows x 512 k = 32 KB of weights + 8 KB of activation planes; 5 stages keep 160 KB of HBM
//     bytes in flight per SM.  One 128-byte swizzled row holds 256 k = two scale groups, so the four K = 64
//     MMAs of a row go to two TMEM accumulators.
// Work decomposition, split-K fix-up, programmatic dependent launch and the fused tensor-parallel all-reduce
// are those of decode_tc.cu.
//
//   * three variants: 2 tokens (default for M <= 2), 4 tokens (converter warps, opt-in) and 8 tokens (opt-in):
//     activations pre-split ONCE per forward by act_presplit_mx4_kernel into six planes of signed base-8 digits
//     and bulk-copied by the producer; the six planes accumulate into ONE set of token columns because plane p's
//     B-side UE8M0 scale factor is 8^p — the block-scaling hardware does the digit recombination.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4-7 = epilogue (TMEM lanes
// 32*(w-4) .. +31), 8-15 = activation converters (idle in the pre-split variant).
#include <cuda.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "act_split.cuh"
#include "gemv_common.cuh"
#include "glu.cuh"
#include "norm.cuh"
#include "sm100.cuh"
#include "mx4_common.cuh"

namespace milab200 {
using namespace gemv;
using namespace sm100;
using namespace mx4;
namespace {

constexpr int kTileRows = 128;
constexpr int kRowK = 256;                          // k per packed 128-byte shared-memory row
constexpr int kGroupK = 128;                        // one PerGroupFp4<128> scale group
constexpr int kPlanes = 8;                          // 2-bit digits of a 16-bit magnitude (in-kernel converter variants)
constexpr int kPlanesPs = 6;                        // signed base-8 digits (pre-split 8-token variant)
constexpr int kDigitBias = 3 * ((1 << 18) - 1) / 7; // 0o333333: adding it turns signed digits -3..4 into octal digits 0..7
constexpr int kARow = kTileRows * 128;              // 16 KB: 128 weight rows x 256 k, packed
constexpr int kTmemCols = 512;
constexpr int kConvWarps = 8;
constexpr int kMxThreads = (8 + kConvWarps) * 32;
constexpr int kXsRing = 64;                         // activation block scales, one entry per group
constexpr int kWsRegions = 4;
constexpr int kMaxSplitItems = 1024;
constexpr int kMaxTiles = 4096;
constexpr int kMaxTokCap = 8;
constexpr int kWsSlotFloats = kMaxTokCap * kTileRows;

// TOKCAP = 2 (M <= 2): a unit is 128 rows x 1024 k — 512 contiguous bytes per weight row, the run length the
// HBM access-pattern probe needs for ~95 % of a linear read (profiles/r1_bw_probe_README.md: 256-byte runs
// 4.4 TB/s, 512-byte runs 5.3 TB/s) — 3 stages, 192 KB of HBM bytes in flight per SM.
// TOKCAP = 4 (M = 3, 4): 512-k units (the plane rows are twice as large), 5 stages, 160 KB in flight.
// TOKCAP = 8 (M = 3..8): activations arrive PRE-SPLIT (act_presplit_mx4_kernel, once per forward) as six planes of
// signed base-8 digits per token, and the producer bulk-copies them next to the weights: no converter warps.  The
// planes of a token do NOT get their own accumulator columns: TMEM reads run at 64 B/clk per SM, and 128 rows x
// 48 columns x 4 B per 8 KB scale group would make the epilogue's tcgen05.ld the bottleneck (measured: 22.6 us on
// Gemma gate_up at M = 4 against 16.2 us for decode_tc.cu).  Instead plane p is its own MMA (N = 16: tokens 0-7,
// columns 8-15 unused) whose B-side UE8M0 scale factor is 2^(3p) = 8^p, all six accumulating into the SAME 16
// columns: the hardware block scaling does the base-8 recombination and a group costs 16 columns of TMEM reads
// whatever M is.  512-k units, 5 stages (160 KB of HBM bytes in flight per SM).
template <int TOKCAP> struct MxShape {
    static constexpr bool kPreSplit = (TOKCAP == 8);
    static constexpr int kNPlanes = kPreSplit ? kPlanesPs : kPlanes;
    static constexpr int kNCols = kPreSplit ? 32 : kNPlanes * TOKCAP;     // accumulator columns per group: 16 / 32 / 32 (24 read)
    static constexpr int kRowsPerUnit = (TOKCAP == 2) ? 4 : 2;            // packed 256-k rows per stage
    static constexpr int kGroupsPerUnit = kRowsPerUnit * 2;               // scale groups per unit: 8 / 4
    static constexpr int kBRow = kNPlanes * TOKCAP * 128;                 // plane rows x 256 k, packed: 2 / 4 / 6 KB
    static constexpr int kAStage = kRowsPerUnit * kARow, kBStage = kRowsPerUnit * kBRow;
    static constexpr int kStages = (TOKCAP == 2) ? 3 : 5;
    static constexpr int kTmemUnits = 3;                                  // accumulator ring, 128 columns per unit
    static constexpr int kSfCol = kTmemUnits * kGroupsPerUnit * kNCols;   // 384 / 256: scale-factor columns (16 for A, then B)
    static constexpr int kSfCols = kPreSplit ? 16 + 8 * kPlanesPs : 32;   // pre-split: one 8-column B region per plane
    static constexpr int kScBatch = kPreSplit ? 1 : ((8 / kGroupsPerUnit) > 0 ? (8 / kGroupsPerUnit) : 1);   // units per scale batch
    static constexpr int kScDepth = kPreSplit ? 8 : 16;                   // weight group scales: two batches in a cp.async ring
    static constexpr int kLdGroups = (64 / kNCols) > 0 ? (64 / kNCols) : 1;   // groups per TMEM read batch (<= 64 registers)
    static constexpr size_t kSmem = (size_t)kStages * (kAStage + kBStage) + kXsRing * kMaxTokCap * 4 +
                                    8 * (2 * kStages + 2 * kTmemUnits) + 64 + kScDepth * kTileRows * 4;
    static_assert(kSmem <= 232448, "exceeds 227 KB of shared memory per CTA");
    static_assert(kXsRing >= kGroupsPerUnit * (kStages + kTmemUnits + 2), "activation-scale ring too short");
    static_assert(kSfCol + kSfCols <= kTmemCols && kScDepth >= 2 * kScBatch * kGroupsPerUnit, "ring sizes");
    // Block-scaled instruction descriptor (kind::mxf4): A/B format E2M1 = 1, UE8M0 scale factors, K = 64 dense,
    // both operands K-major, scale-factor ids 0.
    static constexpr uint32_t kIdesc = (1u << 7) | (1u << 10) | ((uint32_t)(kNCols >> 3) << 17) | (1u << 23) |
                                       ((uint32_t)(kTileRows >> 4) << 24);
};

struct MxParams {
    __nv_bfloat16*       y;
    const __nv_bfloat16* x;
    const float*         scales;
    const __nv_bfloat16* bias;
    float*               ws;            // [items][4][128] partial tiles (P > 1 only)
    int*                 counters;      // [tiles] arrival tickets, all zero between launches
    int M, K, N;
    int KB;                             // K / 128 groups
    int KBU;                            // ceil(KB / 4) units
    int tiles, P, items;
    int early_ld;                       // griddepcontrol.launch_dependents before (1) or after (0) the set-up
    int cl;                             // split-K partials meet through distributed shared memory (cluster of P CTAs)
    int glu, H;                         // fused gate|up -> GLU epilogue: kind (glu.cuh) and hidden width; then N = 2 H,
                                        // tiles = H / 128 logical tiles and KBU counts the units of BOTH halves
    TpExchange tp;
    long long* prof;                    // bring-up only (tools/mx8_timeline.py): CTA 0 records per-unit role timestamps [unit][16]
    int pair;                           // pre-split variant: digit planes per MMA (1, 2, 3), see the MMA issuer
    uint8_t* xp;                        // pre-split variant: plane image [256-k rows][48 x 128 B, swizzled]
    float*   xps;                       //   and block scales [groups][kMaxTokCap]
    int coop;                           // 0: pre-pass kernel; 1: the image is produced by THIS launch (converter warps of CTAs
    int ps_rows;                        //   0 .. ps_rows-1, one 256-k row each, then a grid-wide arrival counter); 2: no image —
                                        //   every CTA's converter warps write the planes of its own units straight into the stages
    int* ps_ctr;                        // [0] rows converted, [1] CTAs that have seen all of them (the last one resets both)
    int sc_l2_ahead;                    // units ahead whose group scales the epilogue pulls into L2 (0 = off)
    NormArgs norm;                      // RMSNorm folded into the activation path (norm.cuh; 2-token variant only): x is normalised
                                        //   in the converter warps' registers, bit for bit what RmsNorm.Bf16.cu would have stored
};

#define MX_PROF(slot)                                                                       \
    do { if constexpr (kPS) { if (p.prof && blockIdx.x == 0 && i < 64) p.prof[i * 16 + (slot)] = clock64() - t_start; } } while (0)
#define MX_PROF_CTA(slot)                                                                   \
    do { if constexpr (kPS) { if (p.prof) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.prof[1024 + blockIdx.x * 4 + (slot)] = t_; } } } while (0)

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(pred));
    return pred != 0;
}

struct Cursor {
    // Two decompositions of the tiles x KBU units of a launch, chosen on the host:
    //  * p.P > 0 (items): tiles x P work items, item (tile, j) = the j-th of P equal runs of a tile's units;
    //    CTA c takes items c, c + G, ...
    //  * p.P == 0 (stream-K): the units, tile-major, are cut into G equal contiguous ranges, one per CTA, so every
    //    SM streams the same number of bytes whatever N and K are; a range crosses tile boundaries, the segments
    //    of a tile cut by a range boundary meet in the workspace exactly like the P partials of an item split.
    int it, tile, ub, ub_end;       // items: item index / stream-K: current unit, ub_end = end of this CTA's range
    bool sk;
    __device__ __forceinline__ void load(const MxParams& p)
    {
        if (it < p.items) {
            tile = it / p.P;
            const int j = it - tile * p.P;
            ub = (int)((long long)j * p.KBU / p.P);
            ub_end = (int)((long long)(j + 1) * p.KBU / p.P);
        }
    }
    __device__ __forceinline__ void start(int cta, const MxParams& p)
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
    __device__ __forceinline__ bool valid(const MxParams& p) const { return sk ? (it < ub_end) : (it < p.items); }
    __device__ __forceinline__ bool item_end(const MxParams& p) const
    {
        return sk ? (ub == p.KBU - 1 || it == ub_end - 1) : (ub == ub_end - 1);
    }
    __device__ __forceinline__ void next(const MxParams& p, int G)
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

// ---- activation split of the 8-token variant --------------------------------------------------------
// One packed 256-k row of the plane image: warp t = token t, lane L = the 8 activations at k = 8 L .. 8 L + 7 (lanes
// 0-15 the row's first scale group, 16-31 the second).  With E = exponent of the token's block maximum,
// u = rn(x 2^(15-E)) is a signed 17-bit integer; u + 0o333333 has octal digits o_p, and d_p = o_p - 3 in {-3 .. 4}
// are signed base-8 digits with u = sum_p 8^p d_p.  Every d_p IS an E2M1 number, so plane p (one nibble per k) is an
// exact operand of kind::mxf4; six planes carry the same 16-bit magnitude as the eight 2-bit planes of the converter
// variants.  Output per row: the 48 x 128-byte shared-memory image of the B operand (row 8 p + t = plane p of token
// t, 128B-swizzled: a 1-D bulk copy drops it into a stage) and, per group, kMaxTokCap block scales 2^(E-15).
// Tokens >= M and groups >= KB (padding of the last unit) are written as zeros.
// 8 consecutive BF16 activations (one uint4) scaled by `scale` -> six plane words: nibble i of W[p] = E2M1 code of
// signed base-8 digit p of element i.  Per element: u + 0o333333 (18 bits), its six 3-bit digits spread to the low
// nibbles of two 16-bit halves, two PRMTs as an 8-entry byte table (digit 0..7 -> code of -3..4), odd elements
// shifted into the high nibbles of their even neighbours, then byte gathers (PRMT) build the plane words.
// (Bit logic checked exhaustively against the digit definition by a Python model of these exact operations.)
__device__ __forceinline__ void base8_planes_x8(const uint4& v, float scale, uint32_t (&W)[kPlanesPs])
{
    const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
    uint32_t A[4], B[4];                        // element pair j: code bytes of digits 0-2 / 3-5 (low nibble = even element)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t ra[2], rb[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float xv = h ? bf16hi(w4[j]) : bf16lo(w4[j]);
            const uint32_t o = (uint32_t)(__float2int_rn(xv * scale) + kDigitBias);
            uint32_t t = (o & 0x1FFu) | ((o << 7) & 0x01FF0000u);
            t = (t & 0x00070007u) | ((t & 0x00380038u) << 1) | ((t & 0x01C001C0u) << 2);
            ra[h] = __byte_perm(0x000A0C0Du, 0x06050402u, t);          // codes of -3, -2, -1, 0 | 1, 2, 3, 4
            rb[h] = __byte_perm(0x000A0C0Du, 0x06050402u, t >> 16);
        }
        A[j] = ra[0] | (ra[1] << 4);
        B[j] = rb[0] | (rb[1] << 4);
    }
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const uint32_t sel = (uint32_t)pl | ((4u + pl) << 4);
        W[pl] = __byte_perm(__byte_perm(A[0], A[1], sel), __byte_perm(A[2], A[3], sel), 0x5410);
        W[3 + pl] = __byte_perm(__byte_perm(B[0], B[1], sel), __byte_perm(B[2], B[3], sel), 0x5410);
    }
}

__device__ __forceinline__ void presplit_mx4_row(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ img,
                                                 float* __restrict__ xs, int M, int K, int KB, int kr, int t, int lane)
{
    const int gh = lane >> 4;
    const int kb = kr * 2 + gh;
    const bool live = (t < M && kb < KB);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (live) v = __ldcg(reinterpret_cast<const uint4*>(x + (size_t)t * K + (size_t)kr * kRowK + lane * 8));
    uint32_t am = __vmaxu2(__vmaxu2(v.x & 0x7FFF7FFFu, v.y & 0x7FFF7FFFu), __vmaxu2(v.z & 0x7FFF7FFFu, v.w & 0x7FFF7FFFu));
#pragma unroll
    for (int lvl = 1; lvl < 16; lvl <<= 1) am = __vmaxu2(am, __shfl_xor_sync(0xffffffffu, am, lvl));
    const uint32_t araw = max(am & 0xFFFFu, am >> 16);
    const bool nonfinite = (araw & 0x7F80u) == 0x7F80u;
    const uint32_t amax = min(araw, 0x7F7Fu);
    int e = 0;                                  // block maximum in [2^e, 2^(e+1))
    if (amax != 0) e = max(-110, (int)(amax >> 7) - 127);
    const float scale = __int_as_float((127 + 15 - e) << 23);
    uint32_t W[kPlanesPs];
    base8_planes_x8(v, scale, W);
#pragma unroll
    for (int pl = 0; pl < kPlanesPs; ++pl) {
        const int row = pl * 8 + t;                 // plane-major: plane p of all 8 tokens is one swizzle atom
        *reinterpret_cast<uint32_t*>(img + (size_t)kr * (kPlanesPs * 8 * 128) + (row >> 3) * 1024 + (row & 7) * 128 +
                                     ((((lane >> 2) ^ row) & 7) << 4) + (lane & 3) * 4) = live ? W[pl] : 0u;
    }
    // Inf/NaN poison the token's output, as they would in FP32: NaN block scale
    if ((lane & 15) == 0)
        xs[(size_t)kb * kMaxTokCap + t] = !live ? 0.0f : (nonfinite ? __int_as_float(0x7FC00000) : __int_as_float((127 + e - 15) << 23));
}

template <int TOKCAP>
__global__ void __launch_bounds__(kMxThreads, 1)
decode_mx4_kernel(const __grid_constant__ CUtensorMap tmap_w, const MxParams p)
{
    using Shape = MxShape<TOKCAP>;
    constexpr int kTokCap = TOKCAP;
    constexpr int kNCols = Shape::kNCols, kRowsPerUnit = Shape::kRowsPerUnit, kGroupsPerUnit = Shape::kGroupsPerUnit;
    constexpr int kBRow = Shape::kBRow, kAStage = Shape::kAStage, kBStage = Shape::kBStage;
    constexpr int kStages = Shape::kStages, kTmemUnits = Shape::kTmemUnits, kSfCol = Shape::kSfCol;
    constexpr int kScBatch = Shape::kScBatch, kLdGroups = Shape::kLdGroups, kScDepth = Shape::kScDepth;
    constexpr bool kPS = Shape::kPreSplit;
    constexpr int kNP = Shape::kNPlanes;
    constexpr uint32_t kIdesc = Shape::kIdesc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0) __trap();
    const uint32_t sA = base;
    const uint32_t sB = sA + kStages * kAStage;
    uint8_t* gB = smem_raw + kStages * kAStage;
    float* g_xs = reinterpret_cast<float*>(gB + kStages * kBStage);
    const uint32_t bars = sB + kStages * kBStage + kXsRing * kMaxTokCap * 4;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + kTmemUnits + s); };
    uint8_t* g_misc = gB + kStages * kBStage + kXsRing * kMaxTokCap * 4 + 8 * (2 * kStages + 2 * kTmemUnits);
    uint32_t* g_tmem_base = reinterpret_cast<uint32_t*>(g_misc);
    int* g_flag = reinterpret_cast<int*>(g_misc + 4);
    float* g_scraw = reinterpret_cast<float*>(g_misc + 64);     // [kScDepth][128]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, KB = p.KB;
    const long long t_start = (kPS && p.prof) ? clock64() : 0;
    (void)t_start;
    if (tid == 0) MX_PROF_CTA(0);

    // ---- one-time setup -------------------------------------------------------------------------
    // The TMA producer (warp 0) needs nothing but the mbarriers: it initialises them and starts streaming
    // weights at once; it only ARRIVES at the two set-up barriers the other warps wait on (TMEM allocation,
    // scale-factor fill), so the first weight bytes are in flight while the rest of the CTA sets up.
    if (p.early_ld) griddep_launch_dependents();        // the next kernel may start its weight prefetch
    if (warp == 0) {
        if (lane == 0) {
            // full: the producer's expect_tx arrival + one arrival per packed row from its converter warp
            // (pre-split variant: the producer's weight arrival + its activation arrival)
            for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), (kPS && p.coop != 2) ? 2 : 1 + kRowsPerUnit); mbar_init(empty_bar(s), 1); }
            for (int s = 0; s < kTmemUnits; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 128); }
            if (p.cl) { mbar_init(smem_u32(g_misc + 16), 128 * (p.P - 1)); mbar_init(smem_u32(g_misc + 24), 1); }
            if (kPS) mbar_init(smem_u32(g_misc + 32), 1);            // "the plane image is complete" (cooperative split)
            fence_mbar_init();
            tma_prefetch_desc(&tmap_w);
        }
        __syncwarp();
        asm volatile("bar.arrive 2, %0;" :: "n"(kMxThreads) : "memory");
        asm volatile("bar.arrive 3, %0;" :: "n"(kMxThreads) : "memory");
    } else {
        // plane rows of unused tokens must read as zero; scale slots must be finite.  (Pre-split image modes: the bulk
        // copies fill whole stages and scale slots — dead rows are zero in the image — and may already be landing, since
        // the producer does not wait for this set-up: stages and scale ring must NOT be touched here.)
        if (!kPS || p.coop == 2) {
            for (int i = tid - 32; i < kStages * kBStage / 16; i += kMxThreads - 32)
                reinterpret_cast<uint4*>(gB)[i] = make_uint4(0, 0, 0, 0);
            for (int i = tid - 32; i < kXsRing * kMaxTokCap; i += kMxThreads - 32) g_xs[i] = 0.0f;
        }
        for (int i = tid - 32; i < kScDepth * kTileRows; i += kMxThreads - 32) g_scraw[i] = 0.0f;
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc(smem_u32(g_tmem_base), kTmemCols);
        tcgen05_fence_before();
        bar_sync(2, kMxThreads);
        tcgen05_fence_after();
    }
    const uint32_t tmem_base = (warp == 0) ? 0u : *g_tmem_base;
    if (warp >= 4 && warp < 8) {
        // unit scale factors (UE8M0 2^0 = 0x7F) for A and B: every lane, 32 columns
        const uint32_t ta = tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + kSfCol;
#pragma unroll
        for (int c = 0; c < 32; c += 8) tmem_st_32x32b_x8(ta + c, 0x7F7F7F7Fu);
        if constexpr (kPS) {
            // B-side scale factors of plane p: UE8M0 2^(3p) = 8^p in every byte of its 8-column region
            // (m = p.pair planes per MMA: region q serves planes m q .. m q + m - 1 in columns 0-7 | 8-15 | 16-23; the
            // scale factor of MMA column c lives in TMEM lane c mod 32 of every lane quadrant)
#pragma unroll
            for (int pl = 0; pl < kNP; ++pl) {
                const uint32_t e = 127 + 3 * (p.pair * pl + min(lane >> 3, p.pair - 1));
                tmem_st_32x32b_x8(ta + 16 + 8 * pl, 0x01010101u * min(e, 254u));
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (warp != 0) {
        tcgen05_fence_before();
        bar_sync(3, kMxThreads);
        tcgen05_fence_after();
    }
    if (!p.early_ld) griddep_launch_dependents();

    Cursor cur;
    cur.start(blockIdx.x, p);
    // fused GLU: a logical tile streams its gate rows' units (ub < KBH), then its up rows' (tile + H/128)
    const int KBH = p.glu ? p.KBU / 2 : p.KBU;
    auto kbu_of = [&](int ub) { return (p.glu && ub >= KBH) ? ub - KBH : ub; };
    auto prow_of = [&](int tile_, int ub) { return (tile_ + ((p.glu && ub >= KBH) ? p.H / kTileRows : 0)) * kTileRows; };

    if (warp == 0) {
        // ===== TMA producer: weights never depend on the previous kernel (no griddepcontrol.wait) =====
        const uint64_t policy = l2_policy_evict_first();
        // Pre-split variant: a unit's plane image and block scales are two more bulk copies.  They are the previous
        // kernel's output, so they go behind griddepcontrol.wait — after the weight loads of the first kStages
        // units are in flight.
        Cursor cb = cur;
        int ib = 0, i = 0;
        bool waited = false;
        auto issue_b_upto = [&](int last) {
            if (!waited) {
                if (p.coop) {
                    mbar_wait(smem_u32(g_misc + 32), 0);         // every row of the image has been written (converter warps)
                    asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes -> bulk-copy reads
                } else {
                    griddep_wait();
                }
                waited = true;
            }
            for (; ib <= last; ++ib, cb.next(p, G)) {
                if (elect_one()) {
                    const int sb = ib % kStages;
                    const size_t kr0 = (size_t)kbu_of(cb.ub) * kRowsPerUnit;
                    mbar_arrive_expect_tx(full_bar(sb), kBStage + kGroupsPerUnit * kMaxTokCap * 4);
                    bulk_load_1d(sB + sb * kBStage, p.xp + kr0 * kBRow, kBStage, full_bar(sb));
                    bulk_load_1d(smem_u32(g_xs) + ((ib * kGroupsPerUnit) % kXsRing) * (kMaxTokCap * 4),
                                 p.xps + kr0 * 2 * kMaxTokCap, kGroupsPerUnit * kMaxTokCap * 4, full_bar(sb));
                }
                __syncwarp();
            }
        };
        for (; cur.valid(p); ++i, cur.next(p, G)) {
            const int s = i % kStages, ph = (i / kStages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
                MX_PROF(0);
                mbar_arrive_expect_tx(full_bar(s), kAStage);
#pragma unroll
                for (int rr = 0; rr < kRowsPerUnit; ++rr)       // bytes past the end of a row are zero-filled by TMA
                    tma_load_2d_hint(sA + s * kAStage + rr * kARow, &tmap_w, (kbu_of(cur.ub) * kRowsPerUnit + rr) * 128,
                                     prow_of(cur.tile, cur.ub), full_bar(s), policy);
                MX_PROF(1);
            }
            __syncwarp();
            if (kPS && p.coop != 2 && i + 1 >= kStages) issue_b_upto(i);
            if (lane == 0) MX_PROF(2);
        }
        if (kPS && p.coop != 2 && ib < i) issue_b_upto(i - 1);   // fewer than kStages units in this CTA
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t tsf = tmem_base + kSfCol;
        for (int i = 0; cur.valid(p); ++i, cur.next(p, G)) {
            const int s = i % kStages, ph = (i / kStages) & 1;
            const int slot = i % kTmemUnits, tph = (i / kTmemUnits) & 1;
            mbar_wait(tempty_bar(slot), tph ^ 1);
            if (lane == 0) MX_PROF(6);
            mbar_wait(full_bar(s), ph);
            tcgen05_fence_after();
            if (elect_one()) {
                MX_PROF(7);
                if constexpr (kPS) {
                    // Plane p = B rows 8 p .. 8 p + 7 (one swizzle atom) scaled by 8^p through its B-side scale
                    // factors; every plane accumulates into the SAME 16 columns of its group.  Back-to-back MMAs
                    // into one accumulator serialise on the accumulate latency (~77 clk each at N = 16: measured),
                    // so consecutive MMAs go round the unit's four group accumulators.
                    //  * single mode: N = 16 reads rows 8 p .. 8 p + 15; the upper 8 (the next plane, or whatever
                    //    follows the image) land in the unused columns 8-15 — every E2M1 nibble is a finite number.
                    //  * pair / triple mode (p.pair = 2 / 3): planes m q .. m q + m - 1 are columns 0-7 | 8-15 | 16-23 of ONE
                    //    MMA with per-column scale factors 8^(m q) | 8^(m q + 1) | ..; the epilogue adds columns t,
                    //    8 + t, 16 + t.  A block-scaled MMA costs ~66 clk whatever its N (measured), so a unit's MMA
                    //    time drops from 48 to 24 to 16 of them — under the 1450 clk a 32 KB unit takes at the HBM rate.
                    const int nmma = kNP / p.pair;
                    const uint32_t bstep = 64u * p.pair;          // m KB per MMA in descriptor units of 16 bytes
                    // N = 16 (one or two planes) or 32 (three planes: columns 24-31 unused)
                    const uint32_t idesc = (kIdesc & ~(0x3Fu << 17)) | ((p.pair == 3 ? 4u : 2u) << 17);
                    for (int pl = 0; pl < nmma; ++pl) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                            for (int rr = 0; rr < kRowsPerUnit; ++rr) {
#pragma unroll
                                for (int kh = 0; kh < 2; ++kh) {
                                    const int k = kh * 2 + kk;
                                    const uint64_t adesc = umma_desc_k_sw128(sA + s * kAStage + rr * kARow);
                                    const uint64_t bdesc = umma_desc_k_sw128(sB + s * kBStage + rr * kBRow);
                                    const uint32_t d = tmem_base + (slot * kGroupsPerUnit + rr * 2 + kh) * kNCols;
                                    umma_mxf4(d, adesc + 2 * k, bdesc + bstep * pl + 2 * k, idesc, tsf, tsf + 16 + 8 * pl,
                                              (uint32_t)(kk | (pl > 0)));
                                }
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int rr = 0; rr < kRowsPerUnit; ++rr) {
                        const uint64_t adesc = umma_desc_k_sw128(sA + s * kAStage + rr * kARow);
                        const uint64_t bdesc = umma_desc_k_sw128(sB + s * kBStage + rr * kBRow);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {           // UMMA K = 64 nibbles = 32 bytes; two MMAs per scale group
                            const uint32_t d = tmem_base + (slot * kGroupsPerUnit + rr * 2 + (k >> 1)) * kNCols;
                            umma_mxf4(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, tsf, tsf + 16, k & 1);
                        }
                    }
                }
                MX_PROF(13);
                umma_commit(empty_bar(s));
                umma_commit(tfull_bar(slot));
                MX_PROF(8);
            }
            __syncwarp();
        }
    } else if (warp >= 8 && kPS) {
        // Pre-split activations.  Cooperative mode: the activations are split ONCE per launch, CTA c writing packed
        // rows c, c + G, ... of the plane image to global memory (warp = token); an arrival counter tells every CTA's
        // producer when all rows are there.  One dependent hand-off (griddepcontrol.wait) + one L2 round trip, where a
        // separate pre-pass kernel costs two kernel boundaries (measured: ~7 us of dependency bubble per launch).
        // Counter protocol: [0] counts converting CTAs, [1] counts CTAs that have OBSERVED completion; the last of
        // those resets both — every later launch touches them only behind its own griddepcontrol.wait, i.e. after
        // this grid (and its reset) has completed.  Only CTAs 0 .. min(G, rows)-1 convert: they are the first to be
        // scheduled, so the hand-off does not wait for the last CTA of the grid to start.
        if (p.coop == 2) {
            // In-kernel split: converter warp cw owns packed row cw % 2 of the units i == cw / 2 (mod 4) of this CTA and
            // writes the six planes of every live token straight into the unit's stage — no image, no second kernel,
            // no grid-wide hand-off; the conversion of four units is in flight at any time.
            const int cw = warp - 8;
            const int rr = cw % kRowsPerUnit, ustride = kConvWarps / kRowsPerUnit, ufirst = cw / kRowsPerUnit;
            const int gh = lane >> 4;
            griddep_wait();                                      // x is the previous kernel's output
            for (int q = 0; q < ufirst && cur.valid(p); ++q) cur.next(p, G);
            for (int i = ufirst; cur.valid(p); i += ustride) {
                const int s = i % kStages, ph = (i / kStages) & 1;
                const int kr = kbu_of(cur.ub) * kRowsPerUnit + rr;
                const bool in_k = (kr * 2 + gh) < KB;
                uint4 cx[kTokCap];
#pragma unroll
                for (int t = 0; t < kTokCap; ++t) {
                    cx[t] = make_uint4(0, 0, 0, 0);
                    if (t < p.M && in_k)
                        cx[t] = __ldcg(reinterpret_cast<const uint4*>(p.x + (size_t)t * p.K + (size_t)kr * kRowK + lane * 8));
                }
                uint32_t am[kTokCap];
#pragma unroll
                for (int t = 0; t < kTokCap; ++t)
                    am[t] = __vmaxu2(__vmaxu2(cx[t].x & 0x7FFF7FFFu, cx[t].y & 0x7FFF7FFFu),
                                     __vmaxu2(cx[t].z & 0x7FFF7FFFu, cx[t].w & 0x7FFF7FFFu));
#pragma unroll
                for (int lvl = 1; lvl < 16; lvl <<= 1) {
#pragma unroll
                    for (int t = 0; t < kTokCap; ++t)
                        if (t < p.M) am[t] = __vmaxu2(am[t], __shfl_xor_sync(0xffffffffu, am[t], lvl));
                }
                mbar_wait(empty_bar(s), ph ^ 1);                 // stage free (its previous MMAs retired)
                uint8_t* brow = gB + s * kBStage + rr * kBRow;
                float* xs_slot = g_xs + ((i * kGroupsPerUnit + rr * 2 + gh) % kXsRing) * kMaxTokCap;
#pragma unroll
                for (int t = 0; t < kTokCap; ++t) {
                    if (t < p.M) {                              // warp-uniform; rows of dead tokens stay zero (set-up)
                        const uint32_t araw = max(am[t] & 0xFFFFu, am[t] >> 16);
                        const bool nonfinite = (araw & 0x7F80u) == 0x7F80u;
                        const uint32_t amax = min(araw, 0x7F7Fu);
                        int e = 0;                              // block maximum in [2^e, 2^(e+1))
                        if (amax != 0) e = max(-110, (int)(amax >> 7) - 127);
                        uint32_t W[kPlanesPs];
                        base8_planes_x8(cx[t], __int_as_float((127 + 15 - e) << 23), W);
#pragma unroll
                        for (int pl = 0; pl < kPlanesPs; ++pl)   // row 8 pl + t: plane pl of all tokens is one swizzle atom
                            *reinterpret_cast<uint32_t*>(brow + pl * 1024 + t * 128 + ((((lane >> 2) ^ t) & 7) << 4) + (lane & 3) * 4) = W[pl];
                        if ((lane & 15) == 0)
                            xs_slot[t] = nonfinite ? __int_as_float(0x7FC00000) : __int_as_float((127 + e - 15) << 23);
                    }
                }
                fence_proxy_async_smem();                        // generic writes -> visible to the MMA's async reads
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
#pragma unroll 1
                for (int q = 0; q < ustride && cur.valid(p); ++q) cur.next(p, G);
            }
        } else if (p.coop) {
            griddep_wait();                                      // x is the previous kernel's output
            const int t = warp - 8;
            for (int kr = blockIdx.x; kr < p.ps_rows; kr += G)
                presplit_mx4_row(p.x, p.xp, p.xps, p.M, p.K, KB, kr, t, lane);
            if ((int)blockIdx.x < p.ps_rows) {
                __threadfence();
                bar_sync(4, kConvWarps * 32);
                if (tid == 256) atomicAdd(p.ps_ctr, 1);
            }
            if (tid == 256) {
                const int target = p.ps_rows < G ? p.ps_rows : G;
                const long long t0 = clock64();
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.ps_ctr) : "memory");
                    if (seen < target && clock64() - t0 > 4000000000LL) __trap();     // ~2 s: a converting CTA never ran
                } while (seen < target);
                mbar_arrive(smem_u32(g_misc + 32));
                if (atomicAdd(p.ps_ctr + 1, 1) == G - 1) {       // last CTA to have seen the full image
                    p.ps_ctr[0] = 0; p.ps_ctr[1] = 0;
                    __threadfence();
                }
            }
        }
    } else if (warp >= 8) {
        // ===== activation converters.  Converter warp cw owns packed row cw % kRowsPerUnit of the units
        //       i == cw / kRowsPerUnit (mod kConvWarps / kRowsPerUnit) of this CTA.  Lane L handles, for every token, the 8 activations at k = 8 L .. 8 L + 7 of the
        //       256-k row: lanes 0-15 are the row's first scale group, lanes 16-31 the second; a token's block
        //       maximum is a 16-lane shuffle reduction. =====
        const int cw = warp - 8;
        const int rr = cw % kRowsPerUnit, ustride = kConvWarps / kRowsPerUnit, ufirst = cw / kRowsPerUnit;
        const int gh = lane >> 4;
        griddep_wait();                                         // x is the previous kernel's output
        float* g_rstd = reinterpret_cast<float*>(g_misc + 48);   // [2] reciprocal RMS per token (fused RMSNorm, 2-token variant)
        if (kTokCap == 2 && p.norm.on) {
            // fused RMSNorm: converter warp cw computes the reciprocal RMS of token cw in the reference's own reduction
            // order (norm.cuh); the normalised BF16 activations then exist only in registers
            if (cw < p.M) {
                const float rs = rms_rstd_select(p.norm, p.x + (size_t)cw * p.K, p.K, lane);
                if (lane == 0) g_rstd[cw] = rs;
            }
            asm volatile("bar.sync 5, %0;" :: "n"(kConvWarps * 32) : "memory");
        }
        uint4 nxt[kTokCap];
        uint4 nw8 = make_uint4(0, 0, 0, 0), nb8 = make_uint4(0, 0, 0, 0);      // norm weight / bias of this lane's 8 k
        auto x_load = [&](int ub) {
            const int kb = (ub * kRowsPerUnit + rr) * 2 + gh;
#pragma unroll
            for (int t = 0; t < kTokCap; ++t) {
                nxt[t] = make_uint4(0, 0, 0, 0);
                if (t < p.M && kb < KB)
                    nxt[t] = __ldcg(reinterpret_cast<const uint4*>(p.x + (size_t)t * p.K + (size_t)(ub * kRowsPerUnit + rr) * kRowK + lane * 8));
            }
            if (kTokCap == 2 && p.norm.on && kb < KB) {
                const size_t k0 = (size_t)(ub * kRowsPerUnit + rr) * kRowK + lane * 8;
                if (p.norm.weight) nw8 = __ldg(reinterpret_cast<const uint4*>(p.norm.weight + k0));
                if (p.norm.bias) nb8 = __ldg(reinterpret_cast<const uint4*>(p.norm.bias + k0));
            }
        };
        for (int q = 0; q < ufirst && cur.valid(p); ++q) cur.next(p, G);
        Cursor pre = cur;
        if (pre.valid(p)) x_load(kbu_of(pre.ub));
        for (int i = ufirst; cur.valid(p); i += ustride) {
            const int s = i % kStages, ph = (i / kStages) & 1;
            uint4 cx[kTokCap];
#pragma unroll
            for (int t = 0; t < kTokCap; ++t) cx[t] = nxt[t];
            if (kTokCap == 2 && p.norm.on) {
                const int kbc = (kbu_of(cur.ub) * kRowsPerUnit + rr) * 2 + gh;
#pragma unroll
                for (int t = 0; t < kTokCap; ++t)
                    if (t < p.M && kbc < KB)
                        cx[t] = rms_apply8(cx[t], g_rstd[t], nw8, nb8, p.norm.weight != nullptr, p.norm.bias != nullptr, p.norm.weight_offset);
            }
#pragma unroll 1
            for (int q = 0; q < ustride && pre.valid(p); ++q) pre.next(p, G);
            if (pre.valid(p)) x_load(kbu_of(pre.ub));            // register prefetch of this warp's next unit
            mbar_wait(empty_bar(s), ph ^ 1);                     // stage free (its previous MMAs retired)
            uint8_t* brow = gB + s * kBStage + rr * kBRow;
            float* xs_slot = g_xs + ((i * kGroupsPerUnit + rr * 2 + gh) % kXsRing) * kMaxTokCap;
            uint32_t am[kTokCap];
#pragma unroll
            for (int t = 0; t < kTokCap; ++t)
                am[t] = __vmaxu2(__vmaxu2(cx[t].x & 0x7FFF7FFFu, cx[t].y & 0x7FFF7FFFu),
                                 __vmaxu2(cx[t].z & 0x7FFF7FFFu, cx[t].w & 0x7FFF7FFFu));
#pragma unroll
            for (int lvl = 1; lvl < 16; lvl <<= 1) {
#pragma unroll
                for (int t = 0; t < kTokCap; ++t)
                    if (t < p.M) am[t] = __vmaxu2(am[t], __shfl_xor_sync(0xffffffffu, am[t], lvl));
            }
            auto convert_token = [&](int t) {
                    const uint32_t araw = max(am[t] & 0xFFFFu, am[t] >> 16);
                    const bool nonfinite = (araw & 0x7F80u) == 0x7F80u;
                    const uint32_t amax = min(araw, 0x7F7Fu);
                    int e = 0;                                  // block maximum in [2^e, 2^(e+1))
                    if (amax != 0) e = max(-110, (int)(amax >> 7) - 127);   // <= 127 by construction
                    const float scale = __int_as_float((127 + 15 - e) << 23);
                    const uint32_t w4[4] = { cx[t].x, cx[t].y, cx[t].z, cx[t].w };
                    uint32_t W[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        W[2 * j] = digit_codes(bf16lo(w4[j]), scale);
                        W[2 * j + 1] = digit_codes(bf16hi(w4[j]), scale);
                    }
                    transpose_nibbles_8x8(W);
                    // plane p of token t is B row t*8 + p: one 8-row swizzle atom per token
                    uint8_t* atom = brow + t * 1024 + (lane & 3) * 4;
#pragma unroll
                    for (int pl = 0; pl < kPlanes; ++pl)
                        *reinterpret_cast<uint32_t*>(atom + pl * 128 + ((((lane >> 2) ^ pl) & 7) << 4)) = W[pl];
                    // Inf/NaN poison the token's output, as they would in FP32: NaN block scale
                    if ((lane & 15) == 0)
                        xs_slot[t] = nonfinite ? __int_as_float(0x7FC00000) : __int_as_float((127 + e - 15) << 23);
            };
            if (p.M == kTokCap) {
                // every token slot is live: no warp-uniform guards, so the independent per-token chains interleave
#pragma unroll
                for (int t = 0; t < kTokCap; ++t) convert_token(t);
            } else {
#pragma unroll
                for (int t = 0; t < kTokCap; ++t)
                    if (t < p.M) convert_token(t);
            }
            fence_proxy_async_smem();                            // generic writes -> visible to the MMA's async reads
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
#pragma unroll 1
            for (int q = 0; q < ustride && cur.valid(p); ++q) cur.next(p, G);
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> FP32 promotion -> BF16 rows / split-K fix-up / tensor-parallel exchange =====
        const int r = tid - 128;                                 // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
        griddep_wait();
        float acc[kTokCap];
#pragma unroll
        for (int t = 0; t < kTokCap; ++t) acc[t] = 0.0f;

        // weight group scales through a private cp.async ring, two units (8 groups = one DRAM sector) at a time
        const uint32_t scslot0 = smem_u32(g_scraw) + r * 4;
        Cursor sc = cur;
        auto scale_fetch_batch = [&](int i0) {
#pragma unroll 1
            for (int q = 0; q < kScBatch && sc.valid(p); ++q) {
                const int row = prow_of(sc.tile, sc.ub) + r;
                if (row < p.N) {
                    const float* sp = p.scales + (size_t)row * KB + kbu_of(sc.ub) * kGroupsPerUnit;
#pragma unroll
                    for (int g = 0; g < kGroupsPerUnit; ++g)
                        if (kbu_of(sc.ub) * kGroupsPerUnit + g < KB)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                                         :: "r"(scslot0 + (((i0 + q) * kGroupsPerUnit + g) % kScDepth) * (kTileRows * 4)), "l"(sp + g)
                                         : "memory");
                    // the ring runs one batch ahead and a scale load under a saturated HBM lasts about as long as a unit: pull
                    // the row's scales of the units behind into L2 now so that their cp.async is an L2 hit (decode_chain.cu)
                    if (p.sc_l2_ahead > 0 && !sc.sk) {
                        const int ub2 = sc.ub + p.sc_l2_ahead;
                        const bool up = p.glu && sc.ub >= KBH;
                        if (ub2 < sc.ub_end && (p.glu && ub2 >= KBH) == up && (kbu_of(sc.ub) + p.sc_l2_ahead) * kGroupsPerUnit < KB)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(sp + p.sc_l2_ahead * kGroupsPerUnit) : "memory");
                        if (i0 == 0 && sc.ub + 1 < sc.ub_end && (p.glu && sc.ub + 1 >= KBH) == up && (kbu_of(sc.ub) + 1) * kGroupsPerUnit < KB)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(sp + kGroupsPerUnit) : "memory");
                    }
                }
                sc.next(p, G);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if constexpr (!kPS) scale_fetch_batch(0);
        // Pre-split variant: a unit lasts ~0.7 us at the HBM rate, less than a scale load takes under load, so a
        // one-batch-ahead ring stalls the epilogue on every unit (measured: 2500 clk per unit whatever the MMA and
        // TMEM work).  This thread's (row, group) scales are instead fetched into registers kScAhead units ahead.
        constexpr int kScAhead = 3;
        float scq[kScAhead][kGroupsPerUnit];
        auto scale_fetch_regs = [&](float (&dst)[kGroupsPerUnit]) {
#pragma unroll
            for (int g = 0; g < kGroupsPerUnit; ++g) dst[g] = 0.0f;
            if (sc.valid(p)) {
                const int row = prow_of(sc.tile, sc.ub) + r;
                if (row < p.N) {
                    const float* sp = p.scales + (size_t)row * KB + kbu_of(sc.ub) * kGroupsPerUnit;
#pragma unroll
                    for (int g = 0; g < kGroupsPerUnit; ++g)
                        if (kbu_of(sc.ub) * kGroupsPerUnit + g < KB) dst[g] = __ldg(sp + g);
                }
                sc.next(p, G);
            }
        };
        if constexpr (kPS) {
#pragma unroll
            for (int q = 0; q < kScAhead; ++q) scale_fetch_regs(scq[q]);
        }

        auto store_row = [&](const float (&v)[kTokCap], int tile_) {
            const int row = tile_ * kTileRows + r;
            if (row < p.N) {
                float bv = 0.0f;
                if (p.bias) bv = __bfloat162float(p.bias[row]);
#pragma unroll
                for (int t = 0; t < kTokCap; ++t)
                    if (t < p.M) p.y[(size_t)t * p.N + row] = __float2bfloat16_rn(v[t] + bv);
            }
        };

        // fused row-parallel all-reduce over NVLink peer memory: see decode_tc.cu (same protocol, same buffers)
        auto finish_rows = [&](float (&v)[kTokCap], int tile_) {
            if (p.tp.world <= 1) { store_row(v, tile_); return; }
            const TpExchange& tp = p.tp;
            const int row = tile_ * kTileRows + r;
            uint32_t epoch = 0;                                             // per-row epoch: see decode_tc.cu
            if (row < p.N) { epoch = __ldcg(tp.row_epoch + row) + 1u; __stcg(tp.row_epoch + row, epoch); }
            const size_t slot_w = (size_t)kMaxTok * tp.nmax;
            float sum[kTokCap];
#pragma unroll
            for (int t = 0; t < kTokCap; ++t) sum[t] = 0.0f;
            if (row < p.N) {
                const size_t mine = ((size_t)(epoch & 1u) * tp.world + tp.rank) * slot_w + row;
                for (int q = 0; q < tp.world; ++q) {
                    if (q == tp.rank) continue;
                    uint2* dst = tp.data[q] + mine;
#pragma unroll
                    for (int t = 0; t < kTokCap; ++t)
                        if (t < p.M)
                            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};"
                                         :: "l"(dst + (size_t)t * tp.nmax), "r"(__float_as_uint(v[t])), "r"(epoch) : "memory");
                }
                const long long t0 = clock64();
                for (int q = 0; q < tp.world; ++q) {
                    const uint2* src = tp.data[tp.rank] + ((size_t)(epoch & 1u) * tp.world + q) * slot_w + row;
#pragma unroll
                    for (int t = 0; t < kTokCap; ++t) {
                        if (t < p.M) {
                            if (q == tp.rank) { sum[t] += v[t]; continue; }
                            uint32_t bits, tag;
                            do {
                                asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];"
                                             : "=r"(bits), "=r"(tag) : "l"(src + (size_t)t * tp.nmax) : "memory");
                                if (tag != epoch && clock64() - t0 > 40000000000LL) __trap();
                            } while (tag != epoch);
                            sum[t] += __uint_as_float(bits);
                        }
                    }
                }
            }
            store_row(sum, tile_);
        };

        float gate[kTokCap];                                       // fused GLU: BF16-rounded gate projections of the tile
#pragma unroll
        for (int t = 0; t < kTokCap; ++t) gate[t] = 0.0f;
        const int first_tile = cur.tile;                         // stream-K: tile of this CTA's first segment
        int seg_first_ub = cur.ub;
        for (int i = 0; cur.valid(p); ++i) {
            const int slot = i % kTmemUnits, tph = (i / kTmemUnits) & 1;
            float scn[kGroupsPerUnit];
            if (r == 0) MX_PROF(9);
            if constexpr (kPS) {
                scale_fetch_regs(scn);                           // unit i + kScAhead
            } else if ((i % kScBatch) == 0) {                    // batch i/kScBatch is needed now: fetch the next one
                scale_fetch_batch(i + kScBatch);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
            mbar_wait(tfull_bar(slot), tph);
            if (r == 0) MX_PROF(10);
            tcgen05_fence_after();
            if constexpr (kPS) {
                // column t (+ 8 + t, 16 + t) = token t: the planes were recombined by the block scaling.  24 columns per
                // group are read, two groups at a time (48 registers).
#pragma unroll
                for (int g0 = 0; g0 < kGroupsPerUnit; g0 += 2) {
                    uint32_t da[2][16], db[2][8];
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t ta = tmem_base + lane_base + (slot * kGroupsPerUnit + g0 + g) * kNCols;
                        tmem_ld_32x32b_x16(ta, da[g]);
                        tmem_ld_32x32b_x8(ta + 16, db[g]);
                    }
                    tmem_ld_wait();
                    if (g0 + 2 == kGroupsPerUnit) {              // everything of this unit has been read
                        tcgen05_fence_before();
                        mbar_arrive(tempty_bar(slot));
                        if (r == 0) MX_PROF(11);
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float wsc = scq[0][g0 + g];
                        const float4* xsp = reinterpret_cast<const float4*>(g_xs + ((i * kGroupsPerUnit + g0 + g) % kXsRing) * kMaxTokCap);
                        const float4 xa = xsp[0], xb = xsp[1];
                        const float xv[8] = { xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w };
#pragma unroll
                        for (int t = 0; t < kTokCap; ++t) {
                            float v = __uint_as_float(da[g][t]);
                            if (p.pair >= 2) v += __uint_as_float(da[g][8 + t]);
                            if (p.pair == 3) v += __uint_as_float(db[g][t]);
                            acc[t] = fmaf(v * xv[t], wsc, acc[t]);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < kGroupsPerUnit; ++g) {
#pragma unroll
                    for (int q = 0; q + 1 < kScAhead; ++q) scq[q][g] = scq[q + 1][g];
                    scq[kScAhead - 1][g] = scn[g];
                }
                if (r == 0) MX_PROF(12);
            } else if (kTokCap == 2 && p.M == 1) {
                // one token: only its eight plane columns are read — half the TMEM bytes and ONE read batch per unit
                uint32_t d1[kGroupsPerUnit][kPlanes];
#pragma unroll
                for (int g = 0; g < kGroupsPerUnit; ++g)
                    tmem_ld_32x32b_x8(tmem_base + lane_base + (slot * kGroupsPerUnit + g) * kNCols, d1[g]);
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(tempty_bar(slot));
#pragma unroll
                for (int g = 0; g < kGroupsPerUnit; ++g) {
                    const float wsc = g_scraw[((i * kGroupsPerUnit + g) % kScDepth) * kTileRows + r];
                    const float xs0 = g_xs[((i * kGroupsPerUnit + g) % kXsRing) * kMaxTokCap];
                    float v = __uint_as_float(d1[g][kPlanes - 1]);
#pragma unroll
                    for (int pl = kPlanes - 2; pl >= 0; --pl) v = fmaf(v, 4.0f, __uint_as_float(d1[g][pl]));
                    acc[0] = fmaf(v * xs0, wsc, acc[0]);
                }
            } else {
#pragma unroll
            for (int g0 = 0; g0 < kGroupsPerUnit; g0 += kLdGroups) {
                uint32_t d[kLdGroups][kNCols];
#pragma unroll
                for (int g = 0; g < kLdGroups; ++g) {
                    const uint32_t ta = tmem_base + lane_base + (slot * kGroupsPerUnit + g0 + g) * kNCols;
                    if constexpr (kNCols == 16) tmem_ld_32x32b_x16(ta, d[g]);
                    else if constexpr (kNCols == 32) tmem_ld_32x32b_x32(ta, d[g]);
                }
                tmem_ld_wait();
                if (g0 + kLdGroups == kGroupsPerUnit) {          // everything of this unit has been read
                    tcgen05_fence_before();
                    mbar_arrive(tempty_bar(slot));
                }
#pragma unroll
                for (int g = 0; g < kLdGroups; ++g) {
                    const float wsc = g_scraw[((i * kGroupsPerUnit + g0 + g) % kScDepth) * kTileRows + r];
                    const float4 xs = *reinterpret_cast<const float4*>(g_xs + ((i * kGroupsPerUnit + g0 + g) % kXsRing) * kMaxTokCap);
                    const float xv[4] = { xs.x, xs.y, xs.z, xs.w };
#pragma unroll
                    for (int t = 0; t < kTokCap; ++t) {
                        if (t < p.M) {                          // warp-uniform
                            float v = __uint_as_float(d[g][t * kPlanes + kPlanes - 1]);
#pragma unroll
                            for (int pl = kPlanes - 2; pl >= 0; --pl) v = fmaf(v, 4.0f, __uint_as_float(d[g][t * kPlanes + pl]));
                            acc[t] = fmaf(v * xv[t], wsc, acc[t]);
                        }
                    }
                }
            }
            }

            if (p.glu && cur.ub == KBH - 1) {
                // end of the gate rows: round the gate projection to BF16 exactly as the unfused Linear stores it
                const int grow = cur.tile * kTileRows + r;
                const float rs = 1.0f;
                const float bv = p.bias ? __bfloat162float(p.bias[grow]) : 0.0f;
#pragma unroll
                for (int t = 0; t < kTokCap; ++t) { gate[t] = bf16_round(fmaf(acc[t], rs, bv)); acc[t] = 0.0f; }
            } else if (p.glu && cur.item_end(p)) {
                const int hrow = cur.tile * kTileRows + r, urow = p.H + hrow;
                const float rs = 1.0f;
                const float bv = p.bias ? __bfloat162float(p.bias[urow]) : 0.0f;
#pragma unroll
                for (int t = 0; t < kTokCap; ++t) {
                    if (t < p.M) p.y[(size_t)t * p.H + hrow] = glu_combine(p.glu, gate[t], bf16_round(fmaf(acc[t], rs, bv)));
                    acc[t] = 0.0f;
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
                            for (int t = 0; t < kTokCap; ++t)
                                if (t < p.M) xbuf[t * kTileRows + r] = acc[t];
                            mbar_arrive_release_cluster(mapa_shared(xbar, 0));
                            if (r == 0) mbar_wait(dbar, 0);                 // the leader has read our partial
                            bar_sync(1, 128);
                        } else {
                            mbar_wait_acquire_cluster(xbar, 0);
                            float v[kTokCap];
#pragma unroll
                            for (int t = 0; t < kTokCap; ++t) v[t] = acc[t];
                            for (int j = 1; j < p.P; ++j) {
                                const uint32_t src = mapa_shared(smem_u32(xbuf) + r * 4, j);
#pragma unroll
                                for (int t = 0; t < kTokCap; ++t)
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
                    for (int t = 0; t < kTokCap; ++t)
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
                        float v[kTokCap];
#pragma unroll
                        for (int t = 0; t < kTokCap; ++t) v[t] = 0.0f;
                        const int j0 = cur.sk ? sk_owner(tile * p.KBU, T, G) : 0;
                        const int j1 = cur.sk ? sk_owner((tile + 1) * p.KBU - 1, T, G) : p.P - 1;
                        for (int j = j0; j <= j1; ++j) {
                            const int slot = cur.sk ? 2 * j + (tile != sk_start(j, T, G) / p.KBU ? 1 : 0) : tile * p.P + j;
                            const float* rp = p.ws + (size_t)slot * kWsSlotFloats + r;
#pragma unroll
                            for (int t = 0; t < kTokCap; ++t)
                                if (t < p.M) v[t] += __ldcg(rp + t * kTileRows);
                        }
                        finish_rows(v, tile);
                        if (r == 0) p.counters[tile] = 0;                   // ready for the next launch
                    }
                    }
                }
#pragma unroll
                for (int t = 0; t < kTokCap; ++t) acc[t] = 0.0f;
                seg_first_ub = (cur.ub == p.KBU - 1) ? 0 : cur.ub + 1;
            }
            cur.next(p, G);
        }
    }

    // ---- teardown ----------------------------------------------------------------------------------
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (tid == 0) MX_PROF_CTA(3);
}

#ifdef MILAB200_DIAG
// ---- stand-alone activation pre-pass (MILAB200_MX8_COOP=0): one CTA per packed 256-k row ---------------
__global__ void __launch_bounds__(32 * 8)
act_presplit_mx4_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ img, float* __restrict__ xs,
                        int M, int K, int KB)
{
    griddep_launch_dependents();                // the decode kernel may start streaming its weights
    griddep_wait();                             // x is the previous kernel's output; it may also still read img
    presplit_mx4_row(x, img, xs, M, K, KB, blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31);
}
#endif  // MILAB200_DIAG

// =================================================================================================
// host side
// =================================================================================================
struct MapKey {
    const void* ptr; int N, K;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && N == o.N && K == o.K; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        return reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull ^ ((size_t)k.N << 32) ^ (size_t)k.K;
    }
};

// packed weights as bytes: [N rows, K/2 bytes], box = 128 bytes (256 k) x 128 rows, 128B swizzle
int packed_tensor_map(const void* w, int N, int K, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ w, N, K };
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return MILAB200_E_NO_DEVICE;
    const cuuint64_t dims[2] = { (cuuint64_t)(K / 2), (cuuint64_t)N };
    const cuuint64_t strides[1] = { (cuuint64_t)(K / 2) };
    const cuuint32_t box[2] = { 128u, (cuuint32_t)kTileRows };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MILAB200_E_BAD_SHAPE;
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

struct MxDevice {
    bool ready = false, failed = false;
    int sms = 0;
    float* ws = nullptr;
    int* counters = nullptr;
    uint8_t* ps_img = nullptr;    // kWsRegions x kPsMaxRows x 6 KB: pre-split activation planes (8-token variant)
    float* ps_xs = nullptr;       // kWsRegions x 2 kPsMaxRows x kMaxTokCap block scales
    int* ps_ctr = nullptr;        // kWsRegions x 2 arrival counters of the cooperative split, zero between launches
    std::atomic<unsigned> next_region{0};
};
constexpr int kPsMaxRows = 512;          // K <= 131072 through the pre-split variant
MxDevice g_mx[16];
std::mutex g_mx_mu;

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
// largest M routed here: 0 = off, 2 = default (M <= 2 only), 4 = also the 4-token converter variant for M = 3, 4,
// 8 = the pre-split 8-token variant for M = 3..8.  Both multi-token variants are parity-tested but measured slower
// than decode_tc.cu on one box (Gemma gate_up, us per launch at M = 4 / 8: decode_tc 17.2 / 20.8, 8-token 20.7 / 20.7):
// the 8-token variant's loop runs at 0.88 us per 512-k unit (24 block-scaled MMAs of ~66 clk each on one issue
// thread) and the separate pre-pass kernel adds a ~7 us dependency bubble per launch (tools/mx8_timeline.py).
std::atomic<int> g_mx_max_m{ env_int("MILAB200_DECODE_MX4_MAXM", 2) };

// 8-token variant: activations split by a pre-pass kernel (0), cooperatively inside the decode launch (1, default), or by
// every CTA's converter warps straight into its stages (2: parity-green but latency-bound — Gemma gate_up 28-30 us
// against 20 us for mode 1; with fewer than 8 tokens the warp-uniform guards even serialise the per-token chains)
std::atomic<int> g_mx_coop{ env_int("MILAB200_MX8_COOP", 1) };
std::atomic<int> g_mx_pair{ env_int("MILAB200_MX8_PAIR", 3) };     // 8-token variant: digit planes per MMA (1, 2 or 3)

#ifdef MILAB200_DIAG
#define MX_VARIANT_ATTR(T) (cudaFuncSetAttribute(decode_mx4_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MxShape<T>::kSmem) != cudaSuccess)
#else
#define MX_VARIANT_ATTR(T) false
#endif
MxDevice* mx_device(cudaStream_t stream)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    MxDevice& d = g_mx[dev];
    if (d.ready) return &d;
    if (d.failed) return nullptr;
    std::lock_guard<std::mutex> lk(g_mx_mu);
    if (d.ready) return &d;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (stream && cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cs != cudaStreamCaptureStatusNone) return nullptr;
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10 || d.sms <= 0 || !encode_tiled_fn()) { d.failed = true; return nullptr; }
    const size_t ws_bytes = (size_t)kWsRegions * kMaxSplitItems * kWsSlotFloats * sizeof(float);
    const size_t ct_bytes = (size_t)kWsRegions * kMaxTiles * sizeof(int);
    const size_t pi_bytes = (size_t)kWsRegions * kPsMaxRows * MxShape<8>::kBRow;
    const size_t px_bytes = (size_t)kWsRegions * kPsMaxRows * 2 * kMaxTokCap * sizeof(float);
    if (cudaMalloc(&d.ws, ws_bytes) != cudaSuccess || cudaMalloc(&d.counters, ct_bytes) != cudaSuccess ||
        cudaMalloc(&d.ps_img, pi_bytes) != cudaSuccess || cudaMalloc(&d.ps_xs, px_bytes) != cudaSuccess ||
        cudaMalloc(&d.ps_ctr, kWsRegions * 2 * sizeof(int)) != cudaSuccess ||
        cudaMemset(d.ps_ctr, 0, kWsRegions * 2 * sizeof(int)) != cudaSuccess ||
        cudaMemset(d.counters, 0, ct_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        d.failed = true;
        return nullptr;
    }
    if (cudaFuncSetAttribute(decode_mx4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MxShape<2>::kSmem) != cudaSuccess ||
        MX_VARIANT_ATTR(4) || MX_VARIANT_ATTR(8)) {
        cudaGetLastError();
        d.failed = true;
        return nullptr;
    }
    d.ready = true;
    return &d;
}

int choose_split(int tiles, int KBU, int sms)
{
    static const int forced = env_int("MILAB200_SPLITK", 0);
    int P = 1;
    if (forced > 0) P = forced;
    else if (tiles * 4 < sms * 3) {
        P = sms / tiles;                                   // one wave; items may be a single (1024-k or 512-k) unit:
                                                           // measured 5.95 vs 6.83 us on Gemma o_proj (4096 -> 3840)
    }
    if (P > KBU) P = KBU;
    if (P < 1) P = 1;
    while (P > 1 && tiles * P > kMaxSplitItems) --P;
    return P;
}

}  // namespace

// defined in decode_tc.cu: true once, after a kernel of this library wrote weight storage (no PDL for that launch)
bool tc_take_weights_fresh();
int tc_streamk_mode();
long long* tc_prof_buffer();

// Returns 1 when the shape / device is not eligible (the caller takes decode_tc.cu), else 0 with the launch
// status in *status.
int try_decode_mx4_norm(__nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                        const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status,
                        const TpExchange* tp, int glu, const NormArgs* norm);
int try_decode_mx4(__nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                   const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status,
                   const TpExchange* tp, int glu)
{
    return try_decode_mx4_norm(y, x, w, scales, bias, M, K, N, stream, status, tp, glu, nullptr);
}

int try_decode_mx4_norm(__nv_bfloat16* y, const __nv_bfloat16* x, const uint8_t* w, const float* scales,
                        const __nv_bfloat16* bias, int M, int K, int N, cudaStream_t stream, int* status,
                        const TpExchange* tp, int glu, const NormArgs* norm)
{
    const int maxm = g_mx_max_m.load(std::memory_order_relaxed);
    if (norm && norm->on && M > 2) return 1;                       // fused RMSNorm: 2-token variant only
    if (M < 1 || M > kMaxTokCap || M > maxm || K % kGroupK != 0) return 1;
    const int var = (M <= 2) ? 0 : (maxm <= 4 ? 1 : 2);            // 2-token / 4-token converter / 8-token pre-split variant
#ifndef MILAB200_DIAG
    if (var != 0) return 1;         // the 4- and 8-token variants lose to decode_tc.cu on every routed shape:
                                    // they are compiled only into the diagnostics build (make -C mila_b200/csrc diag)
#endif
    if (var == 1 && M > 4) return 1;
    if ((reinterpret_cast<uintptr_t>(w) & 15) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
    if (glu && (tp || N % (2 * kTileRows) != 0)) return 1;          // fused GLU: see decode_tc.cu
    const int tiles = glu ? N / (2 * kTileRows) : (N + kTileRows - 1) / kTileRows;
    if (tiles > kMaxTiles) return 1;
    MxDevice* d = mx_device(stream);
    if (!d) return 1;
    if (glu && tiles * 4 < d->sms * 3) return 1;
    // M > 8 on a badly unbalanced two-wave shape: the unfused Linear (stream-K) + activation kernel is faster (measured)
    if (glu && M > 8 && tiles > d->sms && (long long)((tiles + d->sms - 1) / d->sms) * d->sms * 5 >= (long long)tiles * 6) return 1;
    CUtensorMap tm;
    if (packed_tensor_map(w, N, K, &tm) != 0) return 1;

    MxParams p;
    p.y = y; p.x = x; p.scales = scales; p.bias = bias;
    p.norm = norm ? *norm : NormArgs();
    static const int sc_l2 = env_int("MILAB200_MX4_SCALE_L2_AHEAD", 3);
    p.sc_l2_ahead = sc_l2;
    const int gpu = (M <= 2) ? MxShape<2>::kGroupsPerUnit : MxShape<4>::kGroupsPerUnit;
    p.M = M; p.K = K; p.N = N; p.KB = K / kGroupK; p.KBU = (p.KB + gpu - 1) / gpu; p.tiles = tiles;
    const bool streamk = !glu && tc_streamk_mode() == 1;          // opt-in only: see decode_tc.cu (slower at M <= 8)
    p.glu = glu; p.H = N / 2;
    if (glu) p.KBU *= 2;
    p.P = streamk ? 0 : (glu ? 1 : choose_split(tiles, p.KBU, d->sms));
    static const int cluster_fixup = env_int("MILAB200_CLUSTER_FIXUP", 1);
    p.cl = (cluster_fixup && !streamk && p.P > 1 && p.P <= 8 && tiles * p.P <= d->sms) ? 1 : 0;
    p.items = streamk ? tiles * p.KBU : tiles * p.P;
    const unsigned region = d->next_region.fetch_add(1) % kWsRegions;
    p.ws = d->ws + (size_t)region * kMaxSplitItems * kWsSlotFloats;
    p.counters = d->counters + (size_t)region * kMaxTiles;
    if (tp) p.tp = *tp; else p.tp.world = 1;
    if (p.tp.world > 1 && (N > p.tp.nmax || tiles > kTpMaxTiles)) return 1;
    static const int early_ld = env_int("MILAB200_EARLY_LD", 0);   // A/B on one box: no gain for this kernel
    p.early_ld = early_ld;
    const int grid = p.items < d->sms ? p.items : d->sms;
    static const int pdl = env_int("MILAB200_PDL", 1);
    bool decode_pdl = !tc_take_weights_fresh();
    p.xp = nullptr; p.xps = nullptr; p.pair = 0; p.coop = 0; p.ps_rows = 0; p.ps_ctr = nullptr;
    p.prof = tc_prof_buffer();
#ifdef MILAB200_DIAG
    if (var == 2) {
        // split the activations once, ahead of the decode kernel
        const int rows = (glu ? p.KBU / 2 : p.KBU) * MxShape<8>::kRowsPerUnit;
        if (rows > kPsMaxRows) return 1;
        uint8_t* img = d->ps_img + (size_t)region * kPsMaxRows * MxShape<8>::kBRow;
        float* pxs = d->ps_xs + (size_t)region * kPsMaxRows * 2 * kMaxTokCap;
        p.coop = g_mx_coop.load(std::memory_order_relaxed);
        p.ps_rows = rows;
        p.ps_ctr = d->ps_ctr + 2 * region;
        if (!p.coop) {
            cudaLaunchConfig_t pc{};
            pc.gridDim = dim3(rows); pc.blockDim = dim3(32 * 8); pc.stream = stream;
            cudaLaunchAttribute pa[1];
            if (pdl && decode_pdl) {
                // (a launch right behind a quantizer keeps stream order: the decode kernel that follows prefetches
                // weights as soon as THIS kernel starts, and this kernel must then start after the quantizer has ended)
                pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                pa[0].val.programmaticStreamSerializationAllowed = 1;
                pc.attrs = pa; pc.numAttrs = 1;
            }
            const cudaError_t pe = cudaLaunchKernelEx(&pc, act_presplit_mx4_kernel, x, img, pxs, M, K, p.KB);
            if (pe != cudaSuccess) { *status = (int)pe; return 0; }
            note_launch("act_presplit_mx4_kernel");
            decode_pdl = true;
        }
        p.xp = img; p.xps = pxs;
        p.pair = g_mx_pair.load(std::memory_order_relaxed);
    }

#endif

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kMxThreads); cfg.stream = stream;
    cfg.dynamicSmemBytes = (var == 0) ? MxShape<2>::kSmem : (var == 1 ? MxShape<4>::kSmem : MxShape<8>::kSmem);
    cudaLaunchAttribute attr[2];
    int nattr = 0;
    if (pdl && decode_pdl) {
        attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[nattr].val.programmaticStreamSerializationAllowed = 1;
        ++nattr;
    }
    if (p.cl) {
        // the P k-splits of a tile as one thread-block cluster: only if that many clusters can be co-resident
        static std::atomic<int> max_clusters[16][3][9];         // per device, per variant, per cluster size
        int dev = 0; cudaGetDevice(&dev);
        attr[nattr].id = cudaLaunchAttributeClusterDimension;
        attr[nattr].val.clusterDim.x = p.P; attr[nattr].val.clusterDim.y = 1; attr[nattr].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = nattr + 1;
        int mc = (dev >= 0 && dev < 16) ? max_clusters[dev][var][p.P].load() : 0;
        if (mc == 0) {
            int n = 0;
            cudaError_t qe = cudaErrorInvalidValue;
            if (var == 0) qe = cudaOccupancyMaxActiveClusters(&n, decode_mx4_kernel<2>, &cfg);
#ifdef MILAB200_DIAG
            else if (var == 1) qe = cudaOccupancyMaxActiveClusters(&n, decode_mx4_kernel<4>, &cfg);
            else qe = cudaOccupancyMaxActiveClusters(&n, decode_mx4_kernel<8>, &cfg);
#endif
            if (qe != cudaSuccess) { cudaGetLastError(); n = -1; }
            mc = n > 0 ? n : -1;
            if (dev >= 0 && dev < 16) max_clusters[dev][var][p.P].store(mc);
        }
        if (mc >= p.tiles) ++nattr; else p.cl = 0;
    }
    cfg.attrs = attr; cfg.numAttrs = nattr;
    cudaError_t e = cudaErrorInvalidValue;
    if (var == 0) e = cudaLaunchKernelEx(&cfg, decode_mx4_kernel<2>, tm, p);
#ifdef MILAB200_DIAG
    else if (var == 1) e = cudaLaunchKernelEx(&cfg, decode_mx4_kernel<4>, tm, p);
    else e = cudaLaunchKernelEx(&cfg, decode_mx4_kernel<8>, tm, p);
#endif
    if (e != cudaSuccess) { *status = (int)e; return 0; }
    note_launch(var == 0 ? "decode_mx4_kernel<fp4g128,packed,t2>" : var == 1 ? "decode_mx4_kernel<fp4g128,packed,t4>"
                                                                             : "decode_mx4_kernel<fp4g128,packed,t8,presplit>");
    *status = 0;
    return 0;
}

void mx4_set_pair(int m) { g_mx_pair.store(m < 1 ? 1 : (m > 3 ? 3 : m)); }
void mx4_set_coop(int mode) { g_mx_coop.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); }
void mx4_set_max_m(int m) { g_mx_max_m.store(m < 0 ? 0 : (m > kMaxTokCap ? kMaxTokCap : m)); }

}  // namespace milab200
