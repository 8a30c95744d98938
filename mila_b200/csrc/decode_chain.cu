// decode_chain.cu — a CHAIN of decode (M = 1..16) Linears as ONE persistent launch.
//
// Why.  A decode Linear streams 8-60 MB of weights; at 7 TB/s that is 1-9 us, and every kernel boundary between two
// dependent Linears costs ~2.8 us during which HBM idles: the last CTA's epilogue -> grid completion -> the next
// kernel's activation load -> convert -> first MMA, with the next kernel's CTAs unable to become resident (and
// prefetch) while this kernel's 227 KB CTAs hold the SMs (profiles/r2j1_ab_tile_rows.txt: 148 balanced CTAs stream
// 25 % fewer bytes each than 112 whole-tile CTAs and take the same 11.3 us).  Mila drives its Linears one forward at a
// time (Gemma.Block.ixx:298-349, Llama.Block.ixx:883) and names CUDA-graph decode as its next lever
// (CHANGELOG.md:252-258); a graph removes host launch cost but not this device-side bubble.  Here the whole list
// of Linears is one kernel, one CTA per SM:
//   * the TMA producer walks the list without ever waiting for a dependency — weights are constants — so while a
//     layer's outputs are still being finished, the SM's stage ring (216 KB) fills with the NEXT layer's weights:
//     148 x 216 KB = 32 MB, 4.5 us of streaming, longer than the dependency chain it hides;
//   * dependencies are device-side: every CTA checks in on done[l] when its epilogue is through with Linear l (stores,
//     bar.sync, red.release); the activation converters of a Linear spin (ld.acquire, one lane per CTA) until the entry
//     it depends on has collected all check-ins — which also means every earlier entry is complete.  Only the
//     converters wait; producer, MMA issuer and epilogue keep their own pace;
//   * every Linear is cut for ONE balanced wave: tiles of R <= 128 rows so that ceil(N / R) [x 2] items cover all
//     SMs with equal bytes (a layer now lasts as long as its busiest SM).  Layers with few rows (down / o_proj) split
//     k in two; the two halves of a tile run on the two CTAs of a thread-block cluster (the kernel is always
//     launched as clusters of 2: 148 = 2 x 74 packs the GPCs exactly) and meet over distributed shared memory;
//   * everything else is decode_tc.cu: weights are the tcgen05 A operand as they lie in HBM (TMA, 128B swizzle),
//     activations are split exactly into two E4M3 planes, every 128-k block is promoted to FP32 from its own TMEM
//     columns, gate|up Linears can apply GeGLU / SwiGLU in the epilogue, row-parallel shards finish their all-reduce
//     over NVLink peer memory in the epilogue (the wait overlaps the next layer's weight stream).
// Results are bit-identical to the per-Linear launches of the same (R, P) arithmetic order and deterministic
// run to run (partials are added in rank order).
//
// Warp roles as in decode_tc.cu: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4-7 = epilogue,
// 8-15 = activation converters.
#include <cuda.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "act_split.cuh"
#include "gemv_common.cuh"
#include "glu.cuh"
#include "mx4_common.cuh"
#include "sm100.cuh"

namespace milab200 {
using namespace gemv;
using namespace sm100;
using namespace mx4;

struct TpContext;
const TpExchange* tp_context_view(const void* ctx, int* nmax);      // tp.cu

namespace {

constexpr int kTileRows = 128;          // UMMA M
constexpr int kBlockK = 128;            // k elements per group == one PerGroupFp4<128> group
constexpr int kXsRing = 32;             // activation-scale ring (>= groups * (stages + tmem units + 2))
constexpr int kABytes = kTileRows * 128;
constexpr int kMaxLayers = 4096;

// One Linear of the chain (device-resident array, read-only during a run).
struct ChainLayer {
    __nv_bfloat16*       y;
    const __nv_bfloat16* x;
    const float*         scales;
    const __nv_bfloat16* bias;
    int M, K, N;
    int KB, KBU;                        // K / 128 groups; unit blocks per item walk (both GLU halves when glu)
    int R, tiles, P, items;             // tile height, ceil(rows / R), k-splits (1 or 2), tiles * P
    int glu, H;                         // fused gate|up -> GLU: kind and hidden width (N = 2 H)
    int dep;                            // chain index whose completion this Linear's activations need (-1: ready at launch)
    int dep_tiles;                      // check-ins that entry collects when complete (= grid size)
    uint32_t a_tx_bytes;                // mbarrier transaction bytes of one weight box
    // tagged activation hand-off between two entries of the chain (M <= 2): the producer's epilogue writes every pair of
    // output features ALSO as one 8-byte word {bf16 pair, run epoch} (`xout`, [M][out / 2]); the consumer's converters poll
    // those words (`xin`, [M][K / 2]) instead of waiting for the producer's check-in count — one store -> load latency per
    // element instead of release + atomic + poll + load per layer, and no waiting for the slowest CTA of the layer
    const uint2* xin;
    uint2*       xout;
    // k split in four (P = 4): the quarters (0, 1) and (2, 3) of a tile are the two CTAs of a cluster each and meet over
    // distributed shared memory as for P = 2; the leader of the second pair then hands its sum to the leader of the first as
    // tagged FP32 words {value, run epoch} through L2 (`qx`, [tiles][M][128 rows]) — the wire format of the tensor-parallel
    // exchange — and the first leader adds in that fixed order: same bits every run
    uint2*       qx;
    // 9..16 tokens: the activations of an entry are split ONCE per run, cooperatively — CTA c converts the 128-k groups
    // c, c + G, ... into the shared-memory image of the B operand (decode_tc.cu act_presplit_kernel's format) in global
    // memory, checks in on img_done, and every CTA's TMA producer bulk-copies the images next to the weights.  In-kernel
    // conversion by every CTA for its own tile costs ~5000 warp-cycles per 16-token group against a 700-cycle unit.
    uint8_t*     img;                   // [KB][32 rows x 128 B] swizzled E4M3 planes (hi rows | lo rows)
    float*       imgxs;                 // [KB][kMaxTok] block scales
    TpExchange tp;
};

// Third operand scheme beside kFp8 / kFp4G128 (E4M3 planes through kind::f8f6f4): FP4 weights kept as PACKED nibbles and fed
// to kind::mxf4 with the activations cut into eight exact 2-bit digit planes — the scheme of decode_mx4.cu (its header
// comment has the derivation), M <= 2.  A byte of HBM is a byte of shared memory, so the stage ring holds 192 KB of weight
// bytes in flight per SM instead of 96 KB for unpacked nibbles.  Unit = 128 rows x 1024 k = 4 packed 256-k rows = 8 groups.
constexpr int kFmtMx4 = 3;
constexpr int kMxPlanes = 8;

template <int FMT, int NCOLS> struct ChShape {
    static constexpr bool kMx = (FMT == kFmtMx4);
    static constexpr bool kCoop = (!kMx && NCOLS == 32);          // 9..16 tokens: cooperative activation split (images)
    static constexpr int HALF = kMx ? 2 : NCOLS / 2;              // token capacity
    static constexpr int kConvWarps = 8;
    static constexpr int kThreads = (8 + kConvWarps) * 32;
    static constexpr int kGroups = kMx ? 8 : ((NCOLS == 16) ? 4 : 2);          // 128-k scale groups per unit
    static constexpr int kAccCols = kMx ? kMxPlanes * 2 : NCOLS;               // accumulator columns per group
    static constexpr int kBBytes = kMx ? 1024 : NCOLS * 128;                   // activation bytes per group (mx: half a 2 KB packed row)
    static constexpr int kAStage = kMx ? 4 * kABytes : kGroups * kABytes;      // weight bytes of one stage (64 KB at NCOLS 16 / mx)
    static constexpr int kBStage = kGroups * kBBytes;
    static constexpr int kStages = kMx ? 3 : ((NCOLS == 16) ? 3 : 5);
    static constexpr int kTmemUnits = kMx ? 3 : 8 / kGroups;
    static constexpr int kSfCol = kTmemUnits * kGroups * kAccCols;             // mx: unit scale factors live behind the accumulators
    static constexpr int kTmemCols = kMx ? 512 : kTmemUnits * kGroups * kAccCols;
    static constexpr int kXsEntry = kMx ? 4 : kMaxTok;                         // floats per activation-scale ring entry
    static constexpr int kXsRingN = kMx ? 64 : kXsRing;
    static constexpr int kScUnits = 2;                            // FP4 group scales: ring of 2 units (this one + the next)
    static constexpr size_t kSmem = (size_t)kStages * (kAStage + kBStage) + kXsRingN * kXsEntry * 4 +
                                    8 * (2 * kStages + 2 * kTmemUnits + 2) + 64 +
                                    kScUnits * kGroups * kTileRows * 4 + HALF * kTileRows * 4;
    static_assert(kXsRingN >= kGroups * (kStages + kTmemUnits + 2), "activation-scale ring too short");
    static_assert(kSmem <= 232448, "exceeds 227 KB of shared memory per CTA");
};


__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// The work of one CTA inside one Linear: items c, c + G, ... of `items` = tiles x P; item (tile, j) = the j-th of P
// equal runs of the tile's unit blocks.
struct Cursor {
    int it, tile, ub, ub_end;
    int items, P, KBU;
    __device__ __forceinline__ void load()
    {
        if (it < items) {
            // P = 1, 2 or 4 (a power of two): item it = (tile, j), the j-th of P equal runs of the tile's units
            const int sh = (P == 4) ? 2 : (P == 2 ? 1 : 0);
            tile = it >> sh;
            const int j = it & (P - 1);
            ub = (j * KBU) >> sh;
            ub_end = ((j + 1) * KBU) >> sh;
        }
    }
    __device__ __forceinline__ void start(int cta, int items_, int P_, int KBU_) { items = items_; P = P_; KBU = KBU_; it = cta; load(); }
    __device__ __forceinline__ bool valid() const { return it < items; }
    __device__ __forceinline__ bool item_end() const { return ub == ub_end - 1; }
    __device__ __forceinline__ void next(int G) { if (++ub == ub_end) { it += G; load(); } }
};

template <int FMT, int NCOLS>
__global__ void __launch_bounds__(ChShape<FMT, NCOLS>::kThreads, 1)
decode_chain_kernel(const CUtensorMap* __restrict__ tmaps, const ChainLayer* __restrict__ layers, const int nlayers,
                    unsigned* __restrict__ done, unsigned* __restrict__ epoch_word, const int la, const int sigmode, const int sc_l2_ahead,
                    const int small_m_epilogue, long long* __restrict__ prof)
{
    // role timeline (tools/chain_timeline.py): prof[(layer * 8 + slot) * gridDim.x + cta] = globaltimer, null in normal runs
    auto stamp = [&](int l_, int slot_) {
        if (prof) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); prof[((size_t)l_ * 8 + slot_) * gridDim.x + blockIdx.x] = t_; }
    };
    using Shape = ChShape<FMT, NCOLS>;
    constexpr bool kIsFp4 = (FMT != kFp8);
    constexpr bool kMx = Shape::kMx;
    constexpr bool kCoop = Shape::kCoop;
    constexpr int HALF = Shape::HALF;
    constexpr int kBBytes = Shape::kBBytes;
    constexpr int NCW = Shape::kConvWarps;
    constexpr int kStages = Shape::kStages;
    constexpr int kGroups = Shape::kGroups, kTmemUnits = Shape::kTmemUnits, kScUnits = Shape::kScUnits;
    constexpr int kAStage = Shape::kAStage, kBStage = Shape::kBStage, kAccCols = Shape::kAccCols;
    constexpr int kXsEntry = Shape::kXsEntry, kXsRingN = Shape::kXsRingN;
    constexpr uint32_t kIdesc = umma_idesc(kIsFp4 ? kFmtE2M1 : kFmtE4M3, kFmtE4M3, kTileRows, NCOLS);
    // kind::mxf4 block-scaled descriptor: A / B E2M1, UE8M0 scale factors, K = 64, N = 16 (decode_mx4.cu MxShape::kIdesc)
    constexpr uint32_t kIdescMx = (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | (1u << 23) | ((uint32_t)(kTileRows >> 4) << 24);
    constexpr uint32_t kTmemCols = Shape::kTmemCols;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0) __trap();
    const uint32_t sA = base;
    const uint32_t sB = sA + kStages * kAStage;
    uint8_t* gB = smem_raw + kStages * kAStage;
    float* g_xs = reinterpret_cast<float*>(gB + kStages * kBStage);
    const uint32_t bars = sB + kStages * kBStage + kXsRingN * kXsEntry * 4;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + kTmemUnits + s); };
    const uint32_t xbar = bars + 8u * (2 * kStages + 2 * kTmemUnits);        // leader: the partner's partial is parked
    const uint32_t dbar = xbar + 8u;                                         // non-leader: the leader has read it
    uint8_t* g_misc = gB + kStages * kBStage + kXsRingN * kXsEntry * 4 + 8 * (2 * kStages + 2 * kTmemUnits + 2);
    uint32_t* g_tmem_base = reinterpret_cast<uint32_t*>(g_misc);
    int* g_ready = reinterpret_cast<int*>(g_misc + 8);                         // highest layer index + 1 whose input is known complete
    float* g_scraw = reinterpret_cast<float*>(g_misc + 64);                  // [kScUnits * kGroups][128] (FP4 only)
    float* g_xbuf = g_scraw + kScUnits * kGroups * kTileRows;               // [HALF][128] leader: the partner's split-K partial lands here

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    // tag of this run's hand-off words: one more than the last completed run's (CTA 0 publishes it at the very end; runs
    // are ordered by the stream, and CTA 0 cannot finish before every CTA has checked in, i.e. long after they read this)
    const uint32_t epoch = __ldcg(epoch_word) + 1u;

    // ---- one-time setup (the producer starts streaming at once, it only ARRIVES at the set-up barrier) ----
    if (warp == 0) {
        if (lane == 0) {
            // full: the producer's expect_tx arrival + one arrival per converter warp that fills a part of the stage's B operand
            for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), kCoop ? 2 : 1 + (kMx ? 4 : kGroups)); mbar_init(empty_bar(s), 1); }
            for (int s = 0; s < kTmemUnits; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 128); }
            mbar_init(xbar, 1); mbar_init(dbar, 1);
            *g_ready = 0;
            fence_mbar_init();
        }
        __syncwarp();
        asm volatile("bar.arrive 2, %0;" :: "n"(Shape::kThreads) : "memory");
        if constexpr (kMx) asm volatile("bar.arrive 3, %0;" :: "n"(Shape::kThreads) : "memory");
    } else {
        if constexpr (!kCoop)           // (image mode: bulk copies fill whole stages — dead rows are zero in the image)
            for (int i = tid - 32; i < kStages * kBStage / 16; i += Shape::kThreads - 32)
                reinterpret_cast<uint4*>(gB)[i] = make_uint4(0, 0, 0, 0);    // unused token rows must read as zero
        if constexpr (kIsFp4)
            for (int i = tid - 32; i < kScUnits * kGroups * kTileRows; i += Shape::kThreads - 32) g_scraw[i] = 0.0f;
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc(smem_u32(g_tmem_base), kTmemCols);
        tcgen05_fence_before();
        bar_sync(2, Shape::kThreads);
        tcgen05_fence_after();
    }
    const uint32_t tmem_base = (warp == 0) ? 0u : *g_tmem_base;
    if constexpr (kMx) {
        // Mila's FP32 group scales are not UE8M0, so the hardware block scaling is neutralised: unit scale factors (0x7F =
        // 2^0) for A and B in every lane, 32 columns; the real scale is applied in the FP32 promotion
        if (warp >= 4 && warp < 8) {
            const uint32_t ta = tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + Shape::kSfCol;
#pragma unroll
            for (int c = 0; c < 32; c += 8) tmem_st_32x32b_x8(ta + c, 0x7F7F7F7Fu);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (warp != 0) {
            tcgen05_fence_before();
            bar_sync(3, Shape::kThreads);
            tcgen05_fence_after();
        }
    }
    // the partner's mbarriers must exist before anybody arrives on them remotely
    cluster_sync_all();

    // Converters: the entry a Linear depends on has collected every CTA's check-in.  Of the warps that convert one unit only
    // one (`poller`) polls the counter in L2 (1184 pollers on one word measurably delayed the release they were waiting
    // for); the others watch a shared-memory word it publishes (highest entry known complete + 1: an entry being complete
    // implies every earlier one is) — they work on the same unit, so that warp always comes by.
    auto wait_entry_complete = [&](int l_, int dep, int dep_tiles, bool poller, bool stamp_it) {
        volatile int* vready = reinterpret_cast<volatile int*>(g_ready);
        if (poller) {
            if (lane == 0) {
                const long long t0 = clock64();
                while (*vready < dep + 1 && ld_acquire_gpu(done + dep) < (unsigned)dep_tiles) {
                    if (clock64() - t0 > 8000000000LL) __trap();                // ~4 s: a CTA of the chain died
                }
                if (*vready < dep + 1) *vready = dep + 1;                       // (a racing smaller value only costs a re-poll)
                if (stamp_it) stamp(l_, 2);
            }
        } else if (lane == 0) {
            const long long t0 = clock64();
            while (*vready < dep + 1) {
                if (clock64() - t0 > 8000000000LL) __trap();
            }
        }
        __syncwarp();
        __threadfence_block();
    };

    if (warp == 0) {
        // ===== TMA producer: walks the whole chain, never waits for a dependency.  While the stage it needs is still
        //       busy (the pipeline is waiting on a dependency, or simply full) it runs a second walker up to `la` units
        //       ahead and pulls those units' boxes from HBM into L2 (cp.async.bulk.prefetch.tensor): the stage ring holds
        //       ~4 us of streaming, the dependency chain between two Linears is ~5 us (tools/chain_timeline.py), and the
        //       units that follow the ring would otherwise be requested in one burst only after the chain resolves =====
        const uint64_t policy = l2_policy_evict_first();
        struct Walker {
            int l, i, KBU, KBH, R, glu, H; uint32_t tx; Cursor cur; bool done;
        };
        auto enter = [&](Walker& w) {                                // first Linear at or after w.l + 1 with work for this CTA
            w.done = true;
            while (++w.l < nlayers) {
                const ChainLayer* L = layers + w.l;
                w.KBU = L->KBU; w.R = L->R; w.glu = L->glu; w.H = L->H; w.tx = L->a_tx_bytes;
                w.KBH = w.glu ? w.KBU / 2 : w.KBU;
                w.cur.start(blockIdx.x, L->items, L->P, w.KBU);
                if (w.cur.valid()) { w.done = false; break; }
            }
        };
        auto step = [&](Walker& w) { ++w.i; w.cur.next(G); if (!w.cur.valid()) enter(w); };
        Walker w, pw;
        w.l = -1; w.i = 0; enter(w);
        pw = w;
        int last_l = -1;
        // image mode: the activation operand of unit ib follows its weights by at most kStages - 1 units; it waits (once
        // per entry) until every CTA has checked in its share of the entry's images
        Walker wb = w;
        int img_ready_l = -1;
        unsigned* img_done = done + nlayers;
        auto issue_b = [&](bool must) -> bool {
            if (wb.done) return false;
            if (img_ready_l < wb.l) {
                const long long t0 = clock64();
                while (ld_acquire_gpu(img_done + wb.l) < (unsigned)G) {
                    if (!must) return false;
                    if (clock64() - t0 > 8000000000LL) __trap();                // ~4 s: a CTA of the chain died
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");         // generic-proxy image writes -> bulk-copy reads
                img_ready_l = wb.l;
            }
            if (elect_one()) {
                const ChainLayer* Lb = layers + wb.l;
                const int sb = wb.i % kStages;
                const int kbu = (wb.glu && wb.cur.ub >= wb.KBH) ? wb.cur.ub - wb.KBH : wb.cur.ub;
                const size_t kb0 = (size_t)kbu * kGroups;
                mbar_arrive_expect_tx(full_bar(sb), kGroups * (kBBytes + kMaxTok * 4));
                bulk_load_1d(sB + sb * kBStage, Lb->img + kb0 * kBBytes, kGroups * kBBytes, full_bar(sb));
                bulk_load_1d(smem_u32(g_xs) + ((wb.i * kGroups) % kXsRingN) * (kXsEntry * 4), Lb->imgxs + kb0 * kMaxTok,
                             kGroups * kMaxTok * 4, full_bar(sb));
            }
            __syncwarp();
            step(wb);
            return true;
        };
        while (!w.done) {
            const int s = w.i % kStages, ph = (w.i / kStages) & 1;
            if (la > 0 && !mbar_try_wait(empty_bar(s), ph ^ 1)) {
                while (!pw.done && pw.i < w.i + la) {
                    if (pw.i >= w.i && elect_one()) {
                        const bool up = pw.glu && pw.cur.ub >= pw.KBH;
                        const int kbu = up ? pw.cur.ub - pw.KBH : pw.cur.ub;
                        const int prow = pw.cur.tile * pw.R + (up ? pw.H : 0);
                        if constexpr (kMx) {
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) tma_prefetch_l2_2d(tmaps + pw.l, (kbu * 4 + rr) * 128, prow);
                        } else {
#pragma unroll
                            for (int g = 0; g < kGroups; ++g) tma_prefetch_l2_2d(tmaps + pw.l, (kbu * kGroups + g) * kBlockK, prow);
                        }
                    }
                    __syncwarp();
                    step(pw);
                }
            }
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
                if (w.l != last_l) stamp(w.l, 0);
                stamp(w.l, 1);
                const bool up = w.glu && w.cur.ub >= w.KBH;
                const int kbu = up ? w.cur.ub - w.KBH : w.cur.ub;
                const int prow = w.cur.tile * w.R + (up ? w.H : 0);
                if constexpr (kMx) {
                    // four packed 256-k rows (128 bytes per weight row each); the tensor map counts bytes
                    mbar_arrive_expect_tx(full_bar(s), 4 * w.tx);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)              // bytes past the end of a row are zero-filled by TMA
                        tma_load_2d_hint(sA + s * kAStage + rr * kABytes, tmaps + w.l, (kbu * 4 + rr) * 128, prow, full_bar(s), policy);
                } else {
                    mbar_arrive_expect_tx(full_bar(s), kGroups * w.tx);
#pragma unroll
                    for (int g = 0; g < kGroups; ++g)           // a group past the end of K is zero-filled by TMA
                        tma_load_2d_hint(sA + s * kAStage + g * kABytes, tmaps + w.l, (kbu * kGroups + g) * kBlockK, prow, full_bar(s), policy);
                }
            }
            last_l = w.l;
            __syncwarp();
            step(w);
            if constexpr (kCoop) {
                while (!wb.done && wb.i + (kStages - 1) < w.i) issue_b(true);   // must not fall further behind
                while (!wb.done && wb.i < w.i && issue_b(false)) { }             // as far ahead as the images allow
            }
        }
        if constexpr (kCoop) { while (!wb.done && wb.i < w.i) issue_b(true); }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int i = 0;
        for (int l = 0; l < nlayers; ++l) {
            const ChainLayer* L = layers + l;
            Cursor cur; cur.start(blockIdx.x, L->items, L->P, L->KBU);
            bool first = true;
            for (; cur.valid(); ++i, cur.next(G)) {
                const int s = i % kStages, ph = (i / kStages) & 1;
                const int slot = i % kTmemUnits, tph = (i / kTmemUnits) & 1;
                mbar_wait(tempty_bar(slot), tph ^ 1);
                mbar_wait(full_bar(s), ph);
                tcgen05_fence_after();
                if (elect_one()) {
                    if (first) stamp(l, 3);
                    stamp(l, 4);
                    first = false;
                    if constexpr (kMx) {
                        const uint32_t tsf = tmem_base + Shape::kSfCol;
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const uint64_t adesc = umma_desc_k_sw128(sA + s * kAStage + rr * kABytes);
                            const uint64_t bdesc = umma_desc_k_sw128(sB + s * kBStage + rr * 2 * kBBytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {       // UMMA K = 64 nibbles = 32 bytes; two MMAs per scale group
                                const uint32_t d = tmem_base + (slot * kGroups + rr * 2 + (k >> 1)) * kAccCols;
                                umma_mxf4(d, adesc + 2 * k, bdesc + 2 * k, kIdescMx, tsf, tsf + 16, k & 1);
                            }
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < kGroups; ++g) {
                            const uint64_t adesc = umma_desc_k_sw128(sA + s * kAStage + g * kABytes);
                            const uint64_t bdesc = umma_desc_k_sw128(sB + s * kBStage + g * kBBytes);
                            const uint32_t d = tmem_base + (slot * kGroups + g) * kAccCols;
#pragma unroll
                            for (int k = 0; k < kBlockK / 32; ++k)
                                umma_f8f6f4(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, k > 0);
                        }
                    }
                    umma_commit(empty_bar(s));
                    umma_commit(tfull_bar(slot));
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8 && kCoop) {
        // ===== 9..16 tokens: cooperative activation split.  Once the entry's input is complete, this CTA converts the 128-k
        //       groups blockIdx, blockIdx + G, ... (warp j = tokens 2j and 2j + 1, lane = (token parity, 8-k segment): the
        //       arithmetic and the bytes of decode_tc.cu's act_presplit_kernel) into the entry's global image and checks in =====
        const int cw = warp - 8;
        const int seg8 = lane & 15, tsub = lane >> 4, m = 2 * cw + tsub;
        unsigned* img_done = done + nlayers;
        for (int l = 0; l < nlayers; ++l) {
            const ChainLayer* L = layers + l;
            const __nv_bfloat16* x = L->x;
            const int M = L->M, K = L->K, KB = L->KB, KBp = ((KB + kGroups - 1) / kGroups) * kGroups;
            if (L->dep >= 0) wait_entry_complete(l, L->dep, L->dep_tiles, cw == 0, cw == 0);
            uint8_t* img = L->img; float* xs = L->imgxs;
            for (int kb = blockIdx.x; kb < KBp; kb += G) {          // (groups KB..KBp-1 pad the last unit: zeros)
                const bool live = (m < M && kb < KB);
                uint4 v = make_uint4(0, 0, 0, 0);
                if (live) v = __ldcg(reinterpret_cast<const uint4*>(x + (size_t)m * K + (size_t)kb * kBlockK + seg8 * 8));
                uint32_t am = __vmaxu2(__vmaxu2(v.x & 0x7FFF7FFFu, v.y & 0x7FFF7FFFu), __vmaxu2(v.z & 0x7FFF7FFFu, v.w & 0x7FFF7FFFu));
#pragma unroll
                for (int lvl = 1; lvl < 16; lvl <<= 1) am = __vmaxu2(am, __shfl_xor_sync(0xffffffffu, am, lvl));
                const uint32_t amax = min(max(am & 0xFFFFu, am >> 16), 0x7F7Fu);
                const int e = (amax != 0) ? max(-100, min(100, (int)(amax >> 7) - 127 - 7)) : 0;
                uint2 hi, lo;
                split_e4m3x8(v, __int_as_float((127 - e) << 23), hi, lo);
                poison_nonfinite(v, hi);                            // no-op for finite values
                uint8_t* row = img + (size_t)kb * kBBytes + (m >> 3) * 1024 + (m & 7) * 128 + ((((seg8 >> 1) ^ (m & 7)) & 7) << 4) + (seg8 & 1) * 8;
                *reinterpret_cast<uint2*>(row) = hi;                // zeros for a dead row: split(0) = (0, 0)
                *reinterpret_cast<uint2*>(row + (HALF >> 3) * 1024) = lo;
                if (seg8 == 0) xs[(size_t)kb * kMaxTok + m] = live ? __int_as_float((127 + e) << 23) : 0.0f;
            }
            // check-in: this CTA's share of the entry's images is written (bar.sync orders the 256 threads' stores before the
            // gpu-scope release of one of them, which is cumulative)
            asm volatile("bar.sync 3, %0;" :: "n"(NCW * 32) : "memory");
            if (cw == 0 && lane == 0) red_release_gpu_add(img_done + l, 1u);
        }
    } else if (warp >= 8 && !kMx) {
        // ===== activation converters: warp cw owns group cw % kGroups of the units i == cw / kGroups (mod ustride) =====
        const int cw = warp - 8;
        const int g = cw % kGroups, ustride = NCW / kGroups, ufirst = cw / kGroups;
        const int seg8 = lane & 15, tsub = lane >> 4;
        constexpr int CH = HALF / 2;
        int i = 0;
        for (int l = 0; l < nlayers; ++l) {
            const ChainLayer* L = layers + l;
            const __nv_bfloat16* x = L->x;
            const int M = L->M, K = L->K, KB = L->KB, KBU = L->KBU;
            const int KBH = L->glu ? KBU / 2 : KBU;
            const int dep = L->dep, dep_tiles = L->dep_tiles;
            Cursor cur; cur.start(blockIdx.x, L->items, L->P, KBU);
            const uint2* xin = L->xin;                              // tagged hand-off (M <= 2): tokens live in chunk j = 0 only
            bool ready = (dep < 0) || (xin != nullptr), have = false;
            uint4 nxt[CH];
            uint4 lla = make_uint4(0, 0, 0, 0), llb = make_uint4(0, 0, 0, 0);      // raw {pair, tag} x 4 of this lane's 8 k
            auto ll_addr = [&](int kb) { return reinterpret_cast<const uint4*>(xin + (size_t)tsub * (K >> 1) + (((size_t)kb * kBlockK + seg8 * 8) >> 1)); };
            auto ll_read = [&](int kb) {
                const uint4* q = ll_addr(kb);
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lla.x), "=r"(lla.y), "=r"(lla.z), "=r"(lla.w) : "l"(q) : "memory");
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(llb.x), "=r"(llb.y), "=r"(llb.z), "=r"(llb.w) : "l"(q + 1) : "memory");
            };
            auto x_load = [&](int ub) {
                const int kb = ((ub >= KBH) ? ub - KBH : ub) * kGroups + g;
                if (xin) {
                    if (tsub < M && kb < KB) ll_read(kb);           // issued early; validated (and re-read) at use
                    return;
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int m = 2 * j + tsub;
                    nxt[j] = make_uint4(0, 0, 0, 0);
                    if (2 * j < M && m < M && kb < KB)
                        nxt[j] = __ldcg(reinterpret_cast<const uint4*>(x + (size_t)m * K + (size_t)kb * kBlockK + seg8 * 8));
                }
            };
            for (; cur.valid(); ++i, cur.next(G)) {
                if (i % ustride != ufirst) continue;
                if (!ready) { wait_entry_complete(l, dep, dep_tiles, g == 0, cw == 0); ready = true; }
                if (!have) x_load(cur.ub);
                uint4 cx[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) cx[j] = nxt[j];
                if (xin) {
                    // every word carries this run's tag once its producer has stored it: spin on the words themselves
                    const int kbc = ((cur.ub >= KBH) ? cur.ub - KBH : cur.ub) * kGroups + g;
#pragma unroll
                    for (int j = 0; j < CH; ++j) cx[j] = make_uint4(0, 0, 0, 0);
                    if (tsub < M && kbc < KB) {
                        const long long t0 = clock64();
                        while (lla.y != epoch || lla.w != epoch || llb.y != epoch || llb.w != epoch) {
                            ll_read(kbc);
                            if (clock64() - t0 > 8000000000LL) __trap();        // ~4 s: the producing CTA died
                        }
                        cx[0] = make_uint4(lla.x, lla.z, llb.x, llb.z);
                    }
                    if (cw == 0 && lane == 0 && !have) stamp(l, 2);
                }
                {   // register prefetch of this warp's next unit inside this Linear
                    Cursor pre = cur;
                    have = true;
#pragma unroll 1
                    for (int q = 0; q < ustride; ++q) { pre.next(G); if (!pre.valid()) { have = false; break; } }
                    if (have) x_load(pre.ub);
                }
                const int s = i % kStages, ph = (i / kStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                uint8_t* bstage = gB + s * kBStage + g * kBBytes;
                float* xs_slot = g_xs + ((i * kGroups + g) % kXsRingN) * kXsEntry;
                uint32_t am[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    am[j] = __vmaxu2(__vmaxu2(cx[j].x & 0x7FFF7FFFu, cx[j].y & 0x7FFF7FFFu),
                                     __vmaxu2(cx[j].z & 0x7FFF7FFFu, cx[j].w & 0x7FFF7FFFu));
#pragma unroll
                for (int lvl = 1; lvl < 16; lvl <<= 1) {
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (M > CH || 2 * j < M) am[j] = __vmaxu2(am[j], __shfl_xor_sync(0xffffffffu, am[j], lvl));
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (M > CH || 2 * j < M) {                       // warp-uniform: chunks past the last token are skipped
                        const int m = 2 * j + tsub;
                        const uint32_t amax = min(max(am[j] & 0xFFFFu, am[j] >> 16), 0x7F7Fu);
                        const int e = (amax != 0) ? max(-100, min(100, (int)(amax >> 7) - 127 - 7)) : 0;
                        uint2 hi, lo;
                        split_e4m3x8(cx[j], __int_as_float((127 - e) << 23), hi, lo);
                        poison_nonfinite(cx[j], hi);                 // no-op for finite values
                        if (m < M) {
                            uint8_t* row = bstage + (m >> 3) * 1024 + (m & 7) * 128 + ((((seg8 >> 1) ^ (m & 7)) & 7) << 4) + (seg8 & 1) * 8;
                            *reinterpret_cast<uint2*>(row) = hi;
                            *reinterpret_cast<uint2*>(row + (HALF >> 3) * 1024) = lo;
                            if (seg8 == 0) xs_slot[m] = __int_as_float((127 + e) << 23);
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
            }
        }
    } else if (warp >= 8) {
        // ===== activation converters, packed-nibble scheme (decode_mx4.cu): warp cw owns packed 256-k row cw % 4 of the units
        //       i == cw / 4 (mod 2).  Lane L handles, for every token, the 8 activations at k = 8 L .. 8 L + 7 of the row
        //       (lanes 0-15: the row's first scale group, 16-31: the second); u = rn(|x| 2^(15-E)) is cut into eight 2-bit
        //       digits, each an exact E2M1 number: plane p of token t is B row 8 t + p =====
        const int cw = warp - 8;
        const int rr = cw & 3, ustride = 2, ufirst = cw >> 2;
        const int gh = lane >> 4;
        int i = 0;
        for (int l = 0; l < nlayers; ++l) {
            const ChainLayer* L = layers + l;
            const __nv_bfloat16* x = L->x;
            const int M = L->M, K = L->K, KB = L->KB, KBU = L->KBU;
            const int KBH = L->glu ? KBU / 2 : KBU;
            const int dep = L->dep, dep_tiles = L->dep_tiles;
            Cursor cur; cur.start(blockIdx.x, L->items, L->P, KBU);
            const uint2* xin = L->xin;
            bool ready = (dep < 0) || (xin != nullptr), have = false;
            uint4 nxt[2];
            uint4 lla[2], llb[2];                                    // tagged hand-off: raw {pair, tag} x 4 per token
            auto krow_of = [&](int ub) { return ((ub >= KBH) ? ub - KBH : ub) * 4 + rr; };      // packed 256-k row index
            auto ll_read = [&](int t, int kr) {
                const uint4* q = reinterpret_cast<const uint4*>(xin + (size_t)t * (K >> 1) + (((size_t)kr * 256 + lane * 8) >> 1));
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lla[t].x), "=r"(lla[t].y), "=r"(lla[t].z), "=r"(lla[t].w) : "l"(q) : "memory");
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(llb[t].x), "=r"(llb[t].y), "=r"(llb[t].z), "=r"(llb[t].w) : "l"(q + 1) : "memory");
            };
            auto x_load = [&](int ub) {
                const int kr = krow_of(ub), kb = kr * 2 + gh;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    nxt[t] = make_uint4(0, 0, 0, 0);
                    if (t < M && kb < KB) {
                        if (xin) ll_read(t, kr);                    // issued early; validated (and re-read) at use
                        else nxt[t] = __ldcg(reinterpret_cast<const uint4*>(x + (size_t)t * K + (size_t)kr * 256 + lane * 8));
                    }
                }
            };
            for (; cur.valid(); ++i, cur.next(G)) {
                if ((i & 1) != ufirst) continue;
                if (!ready) { wait_entry_complete(l, dep, dep_tiles, rr == 0, cw == 0); ready = true; }
                if (!have) x_load(cur.ub);
                uint4 cx[2];
                cx[0] = nxt[0]; cx[1] = nxt[1];
                if (xin) {
                    const int krc = krow_of(cur.ub), kbc = krc * 2 + gh;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        cx[t] = make_uint4(0, 0, 0, 0);
                        if (t < M && kbc < KB) {
                            const long long t0 = clock64();
                            while (lla[t].y != epoch || lla[t].w != epoch || llb[t].y != epoch || llb[t].w != epoch) {
                                ll_read(t, krc);
                                if (clock64() - t0 > 8000000000LL) __trap();    // ~4 s: the producing CTA died
                            }
                            cx[t] = make_uint4(lla[t].x, lla[t].z, llb[t].x, llb[t].z);
                        }
                    }
                    if (cw == 0 && lane == 0 && !have) stamp(l, 2);
                }
                {   // register prefetch of this warp's next unit inside this Linear
                    Cursor pre = cur;
                    have = true;
#pragma unroll 1
                    for (int q = 0; q < ustride; ++q) { pre.next(G); if (!pre.valid()) { have = false; break; } }
                    if (have) x_load(pre.ub);
                }
                const int s = i % kStages, ph = (i / kStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                uint8_t* brow = gB + s * kBStage + rr * 2 * kBBytes;
                float* xs_slot = g_xs + ((i * kGroups + rr * 2 + gh) % kXsRingN) * kXsEntry;
                uint32_t am[2];
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    am[t] = __vmaxu2(__vmaxu2(cx[t].x & 0x7FFF7FFFu, cx[t].y & 0x7FFF7FFFu),
                                     __vmaxu2(cx[t].z & 0x7FFF7FFFu, cx[t].w & 0x7FFF7FFFu));
#pragma unroll
                for (int lvl = 1; lvl < 16; lvl <<= 1) {
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        if (t < M) am[t] = __vmaxu2(am[t], __shfl_xor_sync(0xffffffffu, am[t], lvl));
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t < M) {                                    // warp-uniform
                        const uint32_t araw = max(am[t] & 0xFFFFu, am[t] >> 16);
                        const bool nonfinite = (araw & 0x7F80u) == 0x7F80u;
                        const uint32_t amax = min(araw, 0x7F7Fu);
                        int e = 0;                                  // block maximum in [2^e, 2^(e+1))
                        if (amax != 0) e = max(-110, (int)(amax >> 7) - 127);
                        const float scale = __int_as_float((127 + 15 - e) << 23);
                        const uint32_t w4[4] = { cx[t].x, cx[t].y, cx[t].z, cx[t].w };
                        uint32_t W[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            W[2 * j] = digit_codes(bf16lo(w4[j]), scale);
                            W[2 * j + 1] = digit_codes(bf16hi(w4[j]), scale);
                        }
                        transpose_nibbles_8x8(W);
                        uint8_t* atom = brow + t * 1024 + (lane & 3) * 4;       // one 8-row swizzle atom per token
#pragma unroll
                        for (int pl = 0; pl < kMxPlanes; ++pl)
                            *reinterpret_cast<uint32_t*>(atom + pl * 128 + ((((lane >> 2) ^ pl) & 7) << 4)) = W[pl];
                        // Inf/NaN poison the token's output, as they would in FP32: NaN block scale
                        if ((lane & 15) == 0)
                            xs_slot[t] = nonfinite ? __int_as_float(0x7FC00000) : __int_as_float((127 + e - 15) << 23);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> FP32 promotion -> BF16 rows (+ split-K over DSMEM, GLU, tensor-parallel sum) =====
        const int r = tid - 128;
        const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
        const uint32_t rank = cluster_ctarank();
        uint32_t xph = 0, dph = 0;                               // phases of the split-K hand-off barriers
        bool xbuf_busy = false;                                  // non-leader: the partner may not have read g_xbuf yet
        float acc[HALF];
#pragma unroll
        for (int t = 0; t < HALF; ++t) acc[t] = 0.0f;
        int i = 0;
        for (int l = 0; l < nlayers; ++l) {
            const ChainLayer* L = layers + l;
            __nv_bfloat16* y = L->y;
            const float* scales = L->scales;
            const __nv_bfloat16* bias = L->bias;
            const int M = L->M, N = L->N, KB = L->KB, KBU = L->KBU, R = L->R, P = L->P, glu = L->glu, H = L->H;
            const int KBH = glu ? KBU / 2 : KBU;
            const int tp_world = L->tp.world;
            Cursor cur; cur.start(blockIdx.x, L->items, P, KBU);

            // FP4 group scales: per-thread cp.async ring, one unit ahead inside this Linear
            const uint32_t scslot0 = smem_u32(g_scraw) + r * 4;
            auto scale_fetch = [&](const Cursor& c, int iu, bool entry_start) {
                if constexpr (kIsFp4) {
                    if (c.valid()) {
                        const bool up = glu && c.ub >= KBH;
                        const int kbu = up ? c.ub - KBH : c.ub;
                        const int row = c.tile * R + (up ? H : 0) + r;
                        if (r < R && row < N) {
                            const float* sp = scales + (size_t)row * KB + kbu * kGroups;
#pragma unroll
                            for (int g = 0; g < kGroups; ++g)
                                if (kbu * kGroups + g < KB)
                                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                                                 :: "r"(scslot0 + (((iu % kScUnits) * kGroups + g) * (kTileRows * 4))), "l"(sp + g) : "memory");
                            // The ring is one unit deep ahead and a scale load under a saturated HBM takes about as long as a
                            // unit's epilogue: every unit waited out a DRAM round trip (tools/chain_timeline.py: the epilogue
                            // of the packed-nibble scheme ran ~1.4 us per unit, 5 us behind the last MMA of every entry).  The
                            // row's scales of the units behind are therefore pulled into L2 now (one sector per unit and row;
                            // no registers, no shared memory), so that their cp.async one unit later is an L2 hit.
                            if (sc_l2_ahead > 0) {
                                const int ub2 = c.ub + sc_l2_ahead;
                                if (ub2 < c.ub_end && (glu && ub2 >= KBH) == up && (kbu + sc_l2_ahead) * kGroups < KB)
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(sp + sc_l2_ahead * kGroups) : "memory");
                                if (entry_start && c.ub + 1 < c.ub_end && (glu && c.ub + 1 >= KBH) == up && (kbu + 1) * kGroups < KB)
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(sp + kGroups) : "memory");
                            }
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
            };
            scale_fetch(cur, i, true);

            // per-row constants of the item in flight (FP8 row scale, bias), fetched when the item STARTS: at its end they
            // sit on the dependency chain of the next Linear (a cold miss there cost ~1 us per layer)
            float rs_a = 1.0f, bv_a = 0.0f, rs_b = 1.0f, bv_b = 0.0f;
            auto load_row_consts = [&]() {
                if (!cur.valid()) return;
                const int row0 = cur.tile * R + r;
                const int ra = glu ? min(row0, H - 1) : min(row0, N - 1), rb = H + ra;
                if constexpr (!kIsFp4) { rs_a = __ldg(scales + ra); if (glu) rs_b = __ldg(scales + rb); }
                if (bias) { bv_a = __bfloat162float(bias[ra]); if (glu) bv_b = __bfloat162float(bias[rb]); }
            };
            load_row_consts();

            // Buffer hazards: a consumer that takes its input through tagged words may run ahead of the check-in count of
            // the entry it `depends_on`; before it WRITES anything it therefore makes sure that entry (hence every earlier
            // one) is complete everywhere — polled once, a couple of units into the item, when it has long been true.
            uint2* xout = L->xout;
            const int hz_dep = (L->xin != nullptr) ? L->dep : -1;
            bool hazard_ok = (hz_dep < 0);
            auto hazard_wait = [&]() {
                if (hazard_ok) return;
                if (r == 0) {
                    const long long t0 = clock64();
                    while (ld_acquire_gpu(done + hz_dep) < (unsigned)G) { if (clock64() - t0 > 8000000000LL) __trap(); }
                }
                bar_sync(1, 128);
                hazard_ok = true;
            };
            // one output row of M tokens: plain BF16 store, and the tagged pair words for a chained consumer.  Called by
            // all 128 threads (the pair partner's value comes by shuffle); `live` rows store.
            auto emit_row = [&](const __nv_bfloat16 (&o)[HALF], int row, int width, bool live) {
#pragma unroll
                for (int t = 0; t < HALF; ++t) {
                    if (t < M) {
                        if (live) y[(size_t)t * width + row] = o[t];
                        if (xout) {                                  // (M <= 2 here; R and every tile base are even)
                            const uint32_t mine = live ? (uint32_t)__bfloat16_as_ushort(o[t]) : 0u;
                            const uint32_t other = __shfl_down_sync(0xffffffffu, mine, 1);
                            if (live && !(row & 1))
                                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};"
                                             :: "l"(xout + (size_t)t * (width >> 1) + (row >> 1)), "r"(mine | (other << 16)), "r"(epoch) : "memory");
                        }
                    }
                }
            };
            auto store_row = [&](const float (&v)[HALF], int tile_) {
                const int row = tile_ * R + r;
                const float rs = rs_a, bv = bv_a;
                __nv_bfloat16 o[HALF];
#pragma unroll
                for (int t = 0; t < HALF; ++t) o[t] = __float2bfloat16_rn(fmaf(v[t], rs, bv));
                emit_row(o, row, N, r < R && row < N);
            };
            // row-parallel shard: one-shot all-reduce over NVLink peer memory (protocol: decode_tc.cu finish_rows)
            auto finish_rows = [&](float (&v)[HALF], int tile_) {
                if (tp_world <= 1) { store_row(v, tile_); return; }
                const TpExchange& tp = L->tp;
                const int row = tile_ * R + r;
                const bool live = (r < R && row < N);
                uint32_t epoch = 0;
                if (live) { epoch = __ldcg(tp.row_epoch + row) + 1u; __stcg(tp.row_epoch + row, epoch); }
                const size_t slot_w = (size_t)kMaxTok * tp.nmax;
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
                            if (t < M)
                                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};"
                                             :: "l"(dst + (size_t)t * tp.nmax), "r"(__float_as_uint(v[t])), "r"(epoch) : "memory");
                    }
                    const long long t0 = clock64();
                    for (int q = 0; q < tp.world; ++q) {
                        const uint2* src = tp.data[tp.rank] + ((size_t)(epoch & 1u) * tp.world + q) * slot_w + row;
#pragma unroll
                        for (int t = 0; t < HALF; ++t) {
                            if (t < M) {
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

            float gate[HALF];
#pragma unroll
            for (int t = 0; t < HALF; ++t) gate[t] = 0.0f;
            for (; cur.valid(); ++i) {
                const int slot = i % kTmemUnits, tph = (i / kTmemUnits) & 1;
                if constexpr (kIsFp4) {
                    Cursor nx = cur; nx.next(G);
                    scale_fetch(nx, i + 1, false);               // next unit of this Linear (an empty group past its end)
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                }
                if (!hazard_ok && cur.ub >= cur.ub_end - 2) hazard_wait();       // (ahead of the item's last unit)
                mbar_wait(tfull_bar(slot), tph);
                tcgen05_fence_after();
                if constexpr (kMx) {
                    // 8 groups x 16 columns (token t, plane p at column 8 t + p), four groups per TMEM read batch; the planes
                    // recombine by Horner in base 4, then the block scale 2^(E-15) and the (row, group) weight scale
                    if (M == 1) {
                        // one token: only its eight plane columns are read — half the TMEM bytes, all eight groups in ONE read
                        // batch (the epilogue of this scheme is the chain's critical role: ~1.75 us per unit at two tokens' width)
                        uint32_t d1[kGroups][kMxPlanes];
#pragma unroll
                        for (int g = 0; g < kGroups; ++g) tmem_ld_32x32b_x8(tmem_base + lane_base + (slot * kGroups + g) * kAccCols, d1[g]);
                        tmem_ld_wait();
                        tcgen05_fence_before();
                        mbar_arrive(tempty_bar(slot));
#pragma unroll
                        for (int g = 0; g < kGroups; ++g) {
                            const float wsc = g_scraw[((i % kScUnits) * kGroups + g) * kTileRows + r];
                            const float xs0 = g_xs[((i * kGroups + g) % kXsRingN) * kXsEntry];
                            float v = __uint_as_float(d1[g][kMxPlanes - 1]);
#pragma unroll
                            for (int pl = kMxPlanes - 2; pl >= 0; --pl) v = fmaf(v, 4.0f, __uint_as_float(d1[g][pl]));
                            acc[0] = fmaf(v * xs0, wsc, acc[0]);
                        }
                    } else
#pragma unroll
                    for (int g0 = 0; g0 < kGroups; g0 += 4) {
                        uint32_t d[4][16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) tmem_ld_32x32b_x16(tmem_base + lane_base + (slot * kGroups + g0 + g) * kAccCols, d[g]);
                        tmem_ld_wait();
                        if (g0 + 4 == kGroups) { tcgen05_fence_before(); mbar_arrive(tempty_bar(slot)); }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float wsc = g_scraw[((i % kScUnits) * kGroups + g0 + g) * kTileRows + r];
                            const float4 xs = *reinterpret_cast<const float4*>(g_xs + ((i * kGroups + g0 + g) % kXsRingN) * kXsEntry);
                            const float xv[2] = { xs.x, xs.y };
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                if (t < M) {                        // warp-uniform
                                    float v = __uint_as_float(d[g][t * kMxPlanes + kMxPlanes - 1]);
#pragma unroll
                                    for (int pl = kMxPlanes - 2; pl >= 0; --pl) v = fmaf(v, 4.0f, __uint_as_float(d[g][t * kMxPlanes + pl]));
                                    acc[t] = fmaf(v * xv[t], wsc, acc[t]);
                                }
                            }
                        }
                    }
                } else if (NCOLS == 16 && M <= 2 && small_m_epilogue) {
                    // one or two tokens: only their hi / lo columns are read (2 + 2 of the group's 16) and promoted
                    uint32_t dh[kGroups][2], dl[kGroups][2];
#pragma unroll
                    for (int g = 0; g < kGroups; ++g) {
                        const uint32_t ta = tmem_base + lane_base + (slot * kGroups + g) * kAccCols;
                        tmem_ld_32x32b_x2(ta, dh[g]);
                        tmem_ld_32x32b_x2(ta + HALF, dl[g]);
                    }
                    tmem_ld_wait();
                    tcgen05_fence_before();
                    mbar_arrive(tempty_bar(slot));
#pragma unroll
                    for (int g = 0; g < kGroups; ++g) {
                        float wsc = 1.0f;
                        if constexpr (kIsFp4) wsc = g_scraw[((i % kScUnits) * kGroups + g) * kTileRows + r];
                        const float2 xs = *reinterpret_cast<const float2*>(g_xs + ((i * kGroups + g) % kXsRingN) * kXsEntry);
                        const float xv[2] = { xs.x, xs.y };
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const float dv = fmaf(__uint_as_float(dl[g][t]), 0.0625f, __uint_as_float(dh[g][t]));
                            if constexpr (kIsFp4) acc[t] = fmaf(dv * xv[t], wsc, acc[t]);
                            else                  acc[t] = fmaf(dv, xv[t], acc[t]);
                        }
                    }
                } else {
                uint32_t d[kGroups][NCOLS];
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    const uint32_t ta = tmem_base + lane_base + (slot * kGroups + g) * kAccCols;
                    if constexpr (NCOLS == 16) tmem_ld_32x32b_x16(ta, d[g]);
                    else                       tmem_ld_32x32b_x32(ta, d[g]);
                }
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(tempty_bar(slot));
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    float wsc = 1.0f;
                    if constexpr (kIsFp4) wsc = g_scraw[((i % kScUnits) * kGroups + g) * kTileRows + r];
                    const float4* xs4 = reinterpret_cast<const float4*>(g_xs + ((i * kGroups + g) % kXsRingN) * kXsEntry);
#pragma unroll
                    for (int q = 0; q < (HALF >= 4 ? HALF / 4 : 1); ++q) {
                        const float4 xs = xs4[q];
                        const float xv[4] = { xs.x, xs.y, xs.z, xs.w };
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int t = q * 4 + j;
                            if (t < HALF) {
                                const float dv = fmaf(__uint_as_float(d[g][(HALF + t) % NCOLS]), 0.0625f, __uint_as_float(d[g][t]));
                                if constexpr (kIsFp4) acc[t] = fmaf(dv * xv[j], wsc, acc[t]);
                                else                  acc[t] = fmaf(dv, xv[j], acc[t]);
                            }
                        }
                    }
                }
                }

                if (glu && cur.ub == KBH - 1) {
                    // end of the gate rows: round the gate projection to BF16 exactly as the unfused Linear stores it
#pragma unroll
                    for (int t = 0; t < HALF; ++t) { gate[t] = bf16_round(fmaf(acc[t], rs_a, bv_a)); acc[t] = 0.0f; }
                } else if (glu && cur.item_end()) {
                    const int hrow = cur.tile * R + r;
                    __nv_bfloat16 o[HALF];
                    float up[HALF];
#pragma unroll
                    for (int t = 0; t < HALF; ++t) { up[t] = bf16_round(fmaf(acc[t], rs_b, bv_b)); acc[t] = 0.0f; }
                    glu_combine_many(glu, gate, up, o);          // straight-line: this sits on the hand-off to the next entry
                    hazard_wait();
                    emit_row(o, hrow, H, r < R && hrow < H);
                } else if (cur.item_end()) {
                    const int tile = cur.tile;
                    hazard_wait();
                    if (P == 1) {
                        finish_rows(acc, tile);
                        } else if (rank != 0) {
                        // second k half: push the partial straight into the LEADER's buffer with st.async — the stores
                        // complete transaction bytes on the leader's mbarrier, so no cluster-scope release is needed (a
                        // release.cluster arrive cost ~2 us here while the next Linear's TMA traffic was in flight:
                        // profiles/r2j7_chain_timeline.txt).  Never wait for the leader here, only before pushing again.
                        if (xbuf_busy) { mbar_wait(dbar, dph); dph ^= 1; }
                        const uint32_t rbuf = mapa_shared(smem_u32(g_xbuf) + r * 4, 0), rbar = mapa_shared(xbar, 0);
#pragma unroll
                        for (int t = 0; t < HALF; ++t)
                            if (t < M)
                                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                                             :: "r"(rbuf + t * kTileRows * 4), "r"(__float_as_uint(acc[t])), "r"(rbar) : "memory");
                        if (r == 0) stamp(l, 7);
                        xbuf_busy = true;
                    } else {
                        if (r == 0) { stamp(l, 2); mbar_arrive_expect_tx(xbar, (uint32_t)(kTileRows * M * 4)); }
                        mbar_wait(xbar, xph); xph ^= 1;                         // the partner's 128 x M words have landed here
                        if (r == 0) stamp(l, 7);
                        float v[HALF];
#pragma unroll
                        for (int t = 0; t < HALF; ++t) v[t] = acc[t] + ((t < M) ? g_xbuf[t * kTileRows + r] : 0.0f);
                        bar_sync(1, 128);                                       // every row has been read
                        if (r == 0) mbar_arrive_cluster(mapa_shared(dbar, 1));
                        if (P == 4 && (cur.it & 2)) {
                            // leader of quarters 2, 3: one 8-byte word per (row, token) to the leader of quarters 0, 1
                            uint2* qp = L->qx + (size_t)tile * HALF * kTileRows + r;
#pragma unroll
                            for (int t = 0; t < HALF; ++t)
                                if (t < M)
                                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};"
                                                 :: "l"(qp + t * kTileRows), "r"(__float_as_uint(v[t])), "r"(epoch) : "memory");
                        } else {
                            if (P == 4) {
                                const uint2* qp = L->qx + (size_t)tile * HALF * kTileRows + r;
                                const long long t0 = clock64();
#pragma unroll
                                for (int t = 0; t < HALF; ++t) {
                                    if (t < M) {
                                        uint32_t bits, tag;
                                        do {
                                            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];"
                                                         : "=r"(bits), "=r"(tag) : "l"(qp + t * kTileRows) : "memory");
                                            if (tag != epoch && clock64() - t0 > 8000000000LL) __trap();      // ~4 s: the other pair died
                                        } while (tag != epoch);
                                        v[t] += __uint_as_float(bits);
                                    }
                                }
                            }
                            finish_rows(v, tile);
                        }
                        }
#pragma unroll
                    for (int t = 0; t < HALF; ++t) acc[t] = 0.0f;
                }
                const bool ended = cur.item_end();
                cur.next(G);
                if (ended) load_row_consts();
            }
            if constexpr (kIsFp4) asm volatile("cp.async.wait_group 0;" ::: "memory");
            // check-in: this CTA is through with Linear l (with or without rows of its own in it).  done[l] == gridDim.x
            // therefore means every CTA's epilogue has passed l — l AND every earlier entry are completely stored — which
            // is what makes `depends_on` a safe barrier for buffer reuse.  bar.sync orders the 128 threads' row stores
            // before thread 0's gpu-scope release (cumulative).
            if (r == 0) stamp(l, 6);
            if (sigmode == 1) __threadfence();
            bar_sync(1, 128);
            if (r == 0) {
                if (sigmode == 1) atomicAdd(done + l, 1u); else red_release_gpu_add(done + l, 1u);
                stamp(l, 5);
            }
        }
    }

    // ---- teardown: nobody leaves while its partner may still read its split-K buffer ----
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (blockIdx.x == 0 && tid == 0) __stcg(epoch_word, epoch);      // the next run's words carry epoch + 1
}

// =================================================================================================
// host side
// =================================================================================================

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

struct Chain {
    int device = 0, count = 0, M = 0, fmt = kFp8, ncols = 16, grid = 0;
    bool mx = false;                     // FP4 at M <= 2: packed nibbles through kind::mxf4 (kFmtMx4)
    CUtensorMap* d_tmaps = nullptr;
    ChainLayer* d_layers = nullptr;
    unsigned* d_done = nullptr;          // [count] check-in counters, [count] image check-ins (9..16 tokens), then the run epoch word
    uint8_t* d_img = nullptr;            // 9..16 tokens: per-entry activation images + block scales
    uint2* d_qx = nullptr;               // tagged FP32 partial sums of the four-way k splits (ChainLayer::qx)
    uint2* d_xchg = nullptr;             // tagged hand-off words of every entry that feeds a later one (M <= 2)
    long long* prof = nullptr;          // role-timeline buffer (milab200_chain_set_timeline), normally null
    int l2_lookahead = 0;               // units the producer may pull into L2 ahead of the stage ring
    int sigmode = 0;                    // check-in: 0 = bar.sync + red.release, 1 = per-thread fence + bar.sync + relaxed atomic
    int small_m_epilogue = 1;           // M <= 2 on the E4M3-plane schemes: the epilogue reads only the live tokens' TMEM columns
    int sc_l2_ahead = 3;                // FP4: units ahead whose group scales the epilogue pulls into L2 (0 = off)
    std::vector<ChainLayer> layers;
};

// Tile height and k-splits for ONE balanced wave over `sms` CTAs (see the header comment); split in two (the halves being the
// two CTAs of a cluster) or in four (two such pairs, the second handing its sum to the first through L2, ChainLayer::qx).  cost = waves * units-per-item * (R + c_unit) [+ c_fix] in row-units.
void choose_chain_decomp(int rows, int KBU, int sms, bool allow_split, bool even_rows, int rmin, int max_p, int* R_out, int* P_out)
{
    static const int c_unit = env_int("MILAB200_CHAIN_COST_UNIT", 8), c_fix = env_int("MILAB200_CHAIN_COST_FIXUP", 32);
    static const int forced_p = env_int("MILAB200_CHAIN_SPLITK", 0), forced_r = env_int("MILAB200_CHAIN_TILE_ROWS", 0);
    long long best = -1;
    *R_out = kTileRows; *P_out = 1;
    static const int c_fix4 = env_int("MILAB200_CHAIN_COST_FIXUP4", 96);
    for (int P = 1; P <= (allow_split ? max_p : 1); P *= 2) {
        if (forced_p > 0 && P != forced_p && !(forced_p >= 2 && !allow_split)) continue;
        if (P > 1 && KBU < P) continue;
        const int upi = (KBU + P - 1) / P;
        for (int R = 16; R <= kTileRows; R += (even_rows ? 2 : 1)) {     // (tagged hand-off pairs rows inside a tile)
            if (forced_r > 0 && R != forced_r) continue;
            const long long tiles = (rows + R - 1) / R;
            const long long items = tiles * P;
            if (P >= 2 && items > sms) continue;
            const long long waves = (items + sms - 1) / sms;
            // a unit costs at least what `rmin` rows cost: below that height the fixed per-unit work (MMA issue — a block-scaled
            // kind::mxf4 MMA takes ~66 clk whatever its height —, conversion, TMEM reads) sets the pace, not the bytes
            // (measured: Gemma down 15360 -> 3840 as 26-row tiles: 16 us for 29.5 MB, profiles/r2j19_chain_timeline_gemma.txt)
            const long long cost = waves * upi * ((R > rmin ? R : rmin) + c_unit) + (P == 2 ? c_fix : (P == 4 ? c_fix4 : 0));
            if (best < 0 || cost < best) { best = cost; *R_out = R; *P_out = P; }
        }
    }
}

template <int FMT, int NCOLS>
int launch_chain(const Chain* c, cudaStream_t stream)
{
    constexpr size_t smem = ChShape<FMT, NCOLS>::kSmem;
    static std::atomic<bool> configured[16];
    if (c->device >= 0 && c->device < 16 && !configured[c->device].load()) {
        MILAB200_RETURN_IF_CUDA(cudaFuncSetAttribute(decode_chain_kernel<FMT, NCOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device].store(true);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c->grid); cfg.blockDim = dim3(ChShape<FMT, NCOLS>::kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    long long* prof = c->prof;
    // The CTAs wait on one another (dependency counters), so all of them must be resident at once: one CTA per SM
    // (227 KB of shared memory each) and a grid no larger than the SM count make that so — kernels ahead of this one in any
    // stream finish without needing it, and every wait in the kernel is bounded (trap after seconds, never a hang).
    // MILAB200_CHAIN_COOPERATIVE=1 adds the cooperative-launch attribute so that the driver verifies co-residency; it is
    // off by default because Nsight Compute (12.9) fails cooperative + cluster launches outright ("LaunchFailed", job
    // r2j14), which would make the kernel unprofilable.
    static std::atomic<int> coop_ok[16];                       // 0 unknown, 1 yes, -1 no
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    const int dev = (c->device >= 0 && c->device < 16) ? c->device : 0;
    static const int want_coop = env_int("MILAB200_CHAIN_COOPERATIVE", 0);
    cudaError_t e = cudaErrorNotSupported;
    if (want_coop && coop_ok[dev].load() >= 0) {
        cfg.numAttrs = 2;
        e = cudaLaunchKernelEx(&cfg, decode_chain_kernel<FMT, NCOLS>, (const CUtensorMap*)c->d_tmaps,
                               (const ChainLayer*)c->d_layers, c->count, c->d_done, c->d_done + 2 * c->count, c->l2_lookahead, c->sigmode, c->sc_l2_ahead, c->small_m_epilogue, prof);
        if (e == cudaSuccess) { coop_ok[dev].store(1); return 0; }
        if (coop_ok[dev].load() == 1 || e == cudaErrorCooperativeLaunchTooLarge) return (int)e;      // a real failure
        cudaGetLastError();
        coop_ok[dev].store(-1);
    }
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, decode_chain_kernel<FMT, NCOLS>, (const CUtensorMap*)c->d_tmaps,
                           (const ChainLayer*)c->d_layers, c->count, c->d_done, c->d_done + 2 * c->count, c->l2_lookahead, c->sigmode, c->sc_l2_ahead, c->small_m_epilogue, prof);
    return (int)e;
}

}  // namespace

int weight_tensor_map(const void* w, int N, int K, int fmt, int R, CUtensorMap* out);      // decode_tc.cu

// FP4 weights as PACKED bytes: [N rows, K/2 bytes], box = 128 bytes (256 k) x R rows, 128B swizzle (the operand layout of
// kind::mxf4; decode_mx4.cu packed_tensor_map with a variable box height)
static int packed_nibble_tensor_map(const void* w, int N, int K, int R, CUtensorMap* out)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return MILAB200_E_NO_DEVICE;
    const cuuint64_t dims[2] = { (cuuint64_t)(K / 2), (cuuint64_t)N };
    const cuuint64_t strides[1] = { (cuuint64_t)(K / 2) };
    const cuuint32_t box[2] = { 128u, (cuuint32_t)R };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : MILAB200_E_BAD_SHAPE;
}

}  // namespace milab200

using namespace milab200;

extern "C" {

int milab200_chain_create(const milab200_chain_linear* lin, int count, int outer_size, void** chain_out)
{
    if (!lin || !chain_out || count <= 0 || count > kMaxLayers || outer_size <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (outer_size > kMaxTok) return MILAB200_E_BAD_SHAPE;
    int dev = 0, major = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return MILAB200_E_NO_DEVICE; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10 || sms < 2 || !encode_tiled_fn()) return MILAB200_E_NO_DEVICE;
    static const int max_grid = env_int("MILAB200_CHAIN_GRID", 0);
    int grid = (sms / 2) * 2;
    if (max_grid > 0 && max_grid < grid) grid = (max_grid / 2) * 2;
    if (grid < 2) return MILAB200_E_NO_DEVICE;

    auto* c = new (std::nothrow) Chain();
    if (!c) return MILAB200_E_INVALID_ARGUMENT;
    c->device = dev; c->count = count; c->M = outer_size; c->grid = grid;
    c->ncols = (outer_size <= 8) ? 16 : 32;
    static const int mx_on = env_int("MILAB200_CHAIN_MX4", 1);
    c->mx = (mx_on != 0 && lin[0].group_size == 128 && outer_size <= 2);
    c->sigmode = env_int("MILAB200_CHAIN_SIGNAL", 0);
    c->sc_l2_ahead = env_int("MILAB200_CHAIN_SCALE_L2_AHEAD", 3);
    c->small_m_epilogue = env_int("MILAB200_CHAIN_SMALL_M_EPILOGUE", 1);
    c->l2_lookahead = env_int("MILAB200_CHAIN_L2_LOOKAHEAD", 0);      // measured: 924 tok/s without, 871 / 857 with 4 / 8 units (r2j4)
    const int groups = c->mx ? ChShape<kFmtMx4, 16>::kGroups : ((c->ncols == 16) ? ChShape<kFp8, 16>::kGroups : ChShape<kFp8, 32>::kGroups);
    std::vector<CUtensorMap> tmaps(count);
    c->layers.resize(count);
    static const int ll_on = env_int("MILAB200_CHAIN_HANDOFF", 1);
    const bool use_ll = (ll_on != 0 && outer_size <= 2);          // tagged activation hand-off between entries
    int rc = 0;
    for (int i = 0; i < count && rc == 0; ++i) {
        const milab200_chain_linear& d = lin[i];
        ChainLayer& L = c->layers[i];
        if (!d.out_bf16 || !d.act_bf16 || !d.weight || !d.scales || d.in_features <= 0 || d.out_features <= 0) { rc = MILAB200_E_INVALID_ARGUMENT; break; }
        const int fmt = (d.group_size == 0) ? kFp8 : kFp4G128;
        if (d.group_size != 0 && d.group_size != 128) { rc = (d.group_size == 64) ? MILAB200_E_BAD_SHAPE : MILAB200_E_UNSUPPORTED_GROUP; break; }
        if (i == 0) c->fmt = fmt; else if (fmt != c->fmt) { rc = MILAB200_E_BAD_SHAPE; break; }      // one policy per chain
        const int K = d.in_features, N = d.out_features;
        if (K % kBlockK != 0 || (reinterpret_cast<uintptr_t>(d.weight) & 31) != 0 || (reinterpret_cast<uintptr_t>(d.act_bf16) & 15) != 0) { rc = MILAB200_E_BAD_SHAPE; break; }
        if (d.glu != 0 && (d.glu != kGluGegluTanh && d.glu != kGluSwiglu)) { rc = MILAB200_E_INVALID_ARGUMENT; break; }
        if (d.glu != 0 && (N % 2 != 0 || d.tp_ctx)) { rc = MILAB200_E_BAD_SHAPE; break; }
        if (d.depends_on >= i || d.depends_on < -1) { rc = MILAB200_E_INVALID_ARGUMENT; break; }
        L.y = static_cast<__nv_bfloat16*>(d.out_bf16); L.x = static_cast<const __nv_bfloat16*>(d.act_bf16);
        L.scales = d.scales; L.bias = static_cast<const __nv_bfloat16*>(d.bias_bf16);
        L.M = outer_size; L.K = K; L.N = N; L.KB = K / kBlockK;
        const int KBU1 = (L.KB + groups - 1) / groups;
        const int rows = d.glu ? N / 2 : N;
        int R = kTileRows, P = 1;
        static const int rmin_f8 = env_int("MILAB200_CHAIN_RMIN", 48), rmin_mx = env_int("MILAB200_CHAIN_RMIN_MX4", 96);
        // four-way k splits: by default where a unit's cost is bound by the MMA issue, not its bytes (the packed-nibble scheme);
        // MILAB200_CHAIN_MAX_SPLITK = 2 / 4 overrides for every scheme
        static const int max_p_env = env_int("MILAB200_CHAIN_MAX_SPLITK", 0);
        const int max_p = (max_p_env == 2 || max_p_env == 4) ? max_p_env : (c->mx ? 4 : 2);
        choose_chain_decomp(rows, d.glu ? 2 * KBU1 : KBU1, grid, d.glu == 0, use_ll, c->mx ? rmin_mx : rmin_f8, max_p, &R, &P);
        L.KBU = d.glu ? 2 * KBU1 : KBU1; L.R = R; L.P = P;
        L.tiles = (rows + R - 1) / R; L.items = L.tiles * P;
        L.glu = d.glu; L.H = N / 2;
        L.dep = d.depends_on; L.dep_tiles = grid;       // every CTA checks in on every entry
        L.a_tx_bytes = c->mx ? (uint32_t)(R * 128) : ((fmt == kFp8) ? (uint32_t)(R * kBlockK) : (uint32_t)(R * kBlockK / 2));
        L.xin = nullptr; L.xout = nullptr; L.img = nullptr; L.imgxs = nullptr; L.qx = nullptr;
        L.tp = TpExchange();
        if (d.tp_ctx) {
            int nmax = 0;
            const TpExchange* v = tp_context_view(d.tp_ctx, &nmax);
            if (!v || N > nmax) { rc = MILAB200_E_BAD_SHAPE; break; }
            L.tp = *v;
        }
        if ((c->mx ? packed_nibble_tensor_map(d.weight, N, K, R, &tmaps[i]) : weight_tensor_map(d.weight, N, K, fmt, R, &tmaps[i])) != 0) { rc = MILAB200_E_BAD_SHAPE; break; }
    }
    // tagged hand-off: entry i takes its activations from the latest earlier entry that writes exactly that tensor
    std::vector<size_t> xoff(count, (size_t)-1);
    size_t xwords = 0;
    if (rc == 0 && use_ll) {
        for (int i = 1; i < count; ++i) {
            for (int j = i - 1; j >= 0; --j) {
                if (lin[j].out_bf16 != lin[i].act_bf16) continue;
                const int width = lin[j].glu ? lin[j].out_features / 2 : lin[j].out_features;
                if (width == lin[i].in_features && width % 2 == 0 && c->layers[j].R % 2 == 0) {
                    if (xoff[j] == (size_t)-1) { xoff[j] = xwords; xwords += (size_t)outer_size * (width / 2); }
                    c->layers[i].xin = reinterpret_cast<const uint2*>(xoff[j] + 1);      // patched to an address below
                }
                break;                                              // only the LATEST writer of that tensor counts
            }
        }
    }
    if (rc == 0) {
        cudaError_t e = cudaMalloc(&c->d_tmaps, sizeof(CUtensorMap) * count);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_layers, sizeof(ChainLayer) * count);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_done, sizeof(unsigned) * (2 * count + 1));
        if (e == cudaSuccess && c->ncols == 32 && !c->mx) {
            // images: per entry ceil(KB / groups) * groups groups x (4 KB planes + 64 B scales)
            size_t total = 0;
            std::vector<size_t> off(count);
            for (int i = 0; i < count; ++i) {
                const size_t kbp = (size_t)((c->layers[i].KB + groups - 1) / groups) * groups;
                off[i] = total; total += kbp * (32 * 128 + kMaxTok * 4);
            }
            e = cudaMalloc(&c->d_img, total);
            if (e == cudaSuccess) e = cudaMemset(c->d_img, 0, total);
            for (int i = 0; i < count && e == cudaSuccess; ++i) {
                const size_t kbp = (size_t)((c->layers[i].KB + groups - 1) / groups) * groups;
                c->layers[i].img = c->d_img + off[i];
                c->layers[i].imgxs = reinterpret_cast<float*>(c->d_img + off[i] + kbp * 32 * 128);
            }
        }
        {   // exchange words of the four-way splits
            size_t qwords = 0;
            std::vector<size_t> qoff(count, 0);
            for (int i = 0; i < count; ++i)
                if (c->layers[i].P == 4) { qoff[i] = qwords; qwords += (size_t)c->layers[i].tiles * kMaxTok * kTileRows; }
            if (e == cudaSuccess && qwords) e = cudaMalloc(&c->d_qx, sizeof(uint2) * qwords);
            if (e == cudaSuccess && qwords) e = cudaMemset(c->d_qx, 0, sizeof(uint2) * qwords);
            for (int i = 0; i < count && e == cudaSuccess; ++i)
                if (c->layers[i].P == 4) c->layers[i].qx = c->d_qx + qoff[i];
        }
        if (e == cudaSuccess && xwords) e = cudaMalloc(&c->d_xchg, sizeof(uint2) * xwords);
        if (e == cudaSuccess && xwords) e = cudaMemset(c->d_xchg, 0, sizeof(uint2) * xwords);
        if (e == cudaSuccess) {
            for (int i = 0; i < count; ++i) {
                if (xoff[i] != (size_t)-1) c->layers[i].xout = c->d_xchg + xoff[i];
                if (c->layers[i].xin) c->layers[i].xin = c->d_xchg + (reinterpret_cast<size_t>(c->layers[i].xin) - 1);
            }
        }
        if (e == cudaSuccess) e = cudaMemcpy(c->d_tmaps, tmaps.data(), sizeof(CUtensorMap) * count, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(c->d_layers, c->layers.data(), sizeof(ChainLayer) * count, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(c->d_done, 0, sizeof(unsigned) * (2 * count + 1));
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { cudaGetLastError(); rc = (int)e; }
    }
    if (rc != 0) {
        if (c->d_tmaps) cudaFree(c->d_tmaps);
        if (c->d_layers) cudaFree(c->d_layers);
        if (c->d_done) cudaFree(c->d_done);
        if (c->d_xchg) cudaFree(c->d_xchg);
        if (c->d_qx) cudaFree(c->d_qx);
        if (c->d_img) cudaFree(c->d_img);
        delete c;
        return rc;
    }
    *chain_out = c;
    return 0;
}

int milab200_chain_forward(void* chain, milab200_stream_t stream_)
{
    auto* c = static_cast<Chain*>(chain);
    if (!c) return MILAB200_E_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MILAB200_RETURN_IF_CUDA(cudaMemsetAsync(c->d_done, 0, sizeof(unsigned) * 2 * c->count, stream));
    int rc;
    if (c->fmt == kFp8) rc = (c->ncols == 16) ? launch_chain<kFp8, 16>(c, stream) : launch_chain<kFp8, 32>(c, stream);
    else if (c->mx)     rc = launch_chain<kFmtMx4, 16>(c, stream);
    else                rc = (c->ncols == 16) ? launch_chain<kFp4G128, 16>(c, stream) : launch_chain<kFp4G128, 32>(c, stream);
    if (rc != 0) return rc;
    note_launch(c->fmt == kFp8 ? (c->ncols == 16 ? "decode_chain_kernel<fp8,n16>" : "decode_chain_kernel<fp8,n32>")
                : c->mx ? "decode_chain_kernel<fp4g128,packed,t2>"
                        : (c->ncols == 16 ? "decode_chain_kernel<fp4g128,n16>" : "decode_chain_kernel<fp4g128,n32>"));
    return 0;
}

int milab200_chain_destroy(void* chain)
{
    auto* c = static_cast<Chain*>(chain);
    if (!c) return 0;
    cudaFree(c->d_tmaps); cudaFree(c->d_layers); cudaFree(c->d_done);
    if (c->d_xchg) cudaFree(c->d_xchg);
    if (c->d_qx) cudaFree(c->d_qx);
    if (c->d_img) cudaFree(c->d_img);
    delete c;
    return 0;
}

/* bring-up: device buffer of count * 8 * grid int64 the kernel fills with per-layer role timestamps (null = off) */
int milab200_chain_set_timeline(void* chain, void* buf, int* grid_out)
{
    auto* c = static_cast<Chain*>(chain);
    if (!c) return MILAB200_E_INVALID_ARGUMENT;
    c->prof = static_cast<long long*>(buf);
    if (grid_out) *grid_out = c->grid;
    return 0;
}

/* introspection for tests / tools: tile height and k-splits the chain picked for entry i */
int milab200_chain_describe(void* chain, int index, int* tile_rows, int* ksplits, int* tiles)
{
    auto* c = static_cast<Chain*>(chain);
    if (!c || index < 0 || index >= c->count) return MILAB200_E_INVALID_ARGUMENT;
    if (tile_rows) *tile_rows = c->layers[index].R;
    if (ksplits) *ksplits = c->layers[index].P;
    if (tiles) *tiles = c->layers[index].tiles;
    return 0;
}

}  // extern "C"
