// tma_probe.cu — read-only probes of TMA tile access patterns (development/measurement hook, not a
// product path).  Same question as bw_probe.cu, asked of the tensor-map path the tcgen05 decode
// kernel uses: which (box rows, k-run per stage, stage count, CTA schedule) keeps HBM streaming?
// One producer thread issues 128B-swizzled 2-D tile loads into a shared-memory ring, one consumer
// thread releases each stage as soon as it lands (no math), so the number is the memory system's.
#include <cuda.h>
#include <cstring>

#include "common.cuh"
#include "sm100.cuh"

namespace milab200 {
using namespace sm100;
namespace {

struct TmaProbeParams {
    int rows, slabs;        // matrix: rows x slabs 128-byte (smem) columns
    int R, C, S;            // box rows, slabs per stage, stages
    int mode, L;            // 0 = contiguous range per CTA; 1 = round-robin in chunks of L units
    int hs;                 // handshake: 0 = producer + consumer threads (try_wait), 1 = one thread does both,
                            //            2 = producer + consumer threads polling with test_wait
    const CUtensorMap* tm_global;   // non-null: use this copy of the tensor map in global memory
    long long* prof;                // non-null: CTA 0 writes cycle totals {wait_empty, expect_tx, tma_issue, wait_full, arrive, count, total}
    int slab_elems;         // tensor-map elements per slab (128 for u8 and for 16U4)
    uint32_t tx_bytes;      // per slab box
};

__device__ __forceinline__ void mbar_poll(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

__global__ void __launch_bounds__(64, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm_param, const TmaProbeParams p)
{
    const CUtensorMap& tm = p.tm_global ? *p.tm_global : tm_param;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[64];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x;
    const int S = p.S;
    const uint32_t stage_bytes = (uint32_t)p.C * p.R * 128u;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[32 + s]), 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm);
    }
    __syncthreads();
    // 32-bit, division-free loops: the probe must not be bound by its own address arithmetic
    const int T = (p.rows + p.R - 1) / p.R, CG = p.slabs / p.C;
    const int U = T * CG;
    const int G = gridDim.x, c = blockIdx.x;
    const int u0 = (int)((long long)c * U / G), u1 = (int)((long long)(c + 1) * U / G);
    const int count = u1 - u0;
    const int C = p.C, R = p.R, slab_elems = p.slab_elems;
    const uint32_t tx = p.tx_bytes * p.C;
    if (tid == 0) {
        int tile = u0 / CG, cg = u0 - tile * CG, s = 0, ph = 0;
        long long t_issue = 0, t_wait = 0;
        for (int i = 0; i < count; ++i) {
            const long long t0 = clock64();
            mbar_wait(smem_u32(&bars[32 + s]), ph ^ 1);
            const long long t1 = clock64();
            mbar_arrive_expect_tx(smem_u32(&bars[s]), tx);
            for (int k = 0; k < C; ++k)
                tma_load_2d(base + s * stage_bytes + k * R * 128, &tm, (cg * C + k) * slab_elems, tile * R, smem_u32(&bars[s]));
            t_wait += t1 - t0; t_issue += clock64() - t1;
            if (++cg == CG) { cg = 0; ++tile; }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        if (p.prof && c == 0) { p.prof[0] = t_wait; p.prof[2] = t_issue; p.prof[5] = count; }
    } else if (tid == 32) {
        int s = 0, ph = 0;
        const long long tb = clock64();
        for (int i = 0; i < count; ++i) {
            mbar_wait(smem_u32(&bars[s]), ph);
            mbar_arrive(smem_u32(&bars[32 + s]));
            if (p.prof && c == 0 && i < 120) p.prof[8 + i] = clock64() - tb;
            if (++s == S) { s = 0; ph ^= 1; }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
}  // namespace
}  // namespace milab200

using namespace milab200;

// u4 != 0: the matrix is [rows, row_bytes*2 nibbles] read through 16U4_ALIGN16B (64 global bytes per slab)
extern "C" int milab200_probe_tma(const void* base, int64_t rows, int64_t row_bytes, int u4, int R, int C, int S,
                                       int mode, int L, int promo, int grid, int hs, void* tm_global_buf,
                                       void* prof_buf, milab200_stream_t stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f)
            return MILAB200_E_NO_DEVICE;
        enc = reinterpret_cast<EncodeTiledFn>(f);
    }
    if (R < 8 || R > 256 || C < 1 || S < 1 || S > 32 || (size_t)C * R * 128 * S > 220 * 1024) return MILAB200_E_INVALID_ARGUMENT;
    const int64_t slab_global = u4 ? 64 : 128;
    if (row_bytes % (slab_global * C) != 0) return MILAB200_E_BAD_SHAPE;
    CUtensorMap tm;
    const cuuint64_t dims[2] = { (cuuint64_t)(u4 ? row_bytes * 2 : row_bytes), (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)row_bytes };
    const cuuint32_t box[2] = { 128, (cuuint32_t)R };
    const cuuint32_t estr[2] = { 1, 1 };
    if (enc(&tm, u4 ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base),
            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return MILAB200_E_BAD_SHAPE;
    TmaProbeParams p;
    p.rows = (int)rows; p.slabs = (int)(row_bytes / slab_global); p.R = R; p.C = C; p.S = S; p.mode = mode; p.L = L < 1 ? 1 : L;
    p.slab_elems = 128; p.tx_bytes = (uint32_t)(R * slab_global);
    p.hs = hs; p.tm_global = nullptr; p.prof = static_cast<long long*>(prof_buf);
    if (tm_global_buf) {      // caller-provided 128-byte device buffer (64-byte aligned); plain synchronous copy
        static CUtensorMap last; static void* last_buf = nullptr;
        if (last_buf != tm_global_buf || memcmp(&last, &tm, sizeof(tm)) != 0) {       // not during capture
            MILAB200_RETURN_IF_CUDA(cudaMemcpy(tm_global_buf, &tm, sizeof(tm), cudaMemcpyHostToDevice));
            last = tm; last_buf = tm_global_buf;
        }
        p.tm_global = static_cast<const CUtensorMap*>(tm_global_buf);
    }
    const size_t smem = (size_t)C * R * 128 * S + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        MILAB200_RETURN_IF_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    tma_probe_kernel<<<grid, 64, smem, stream>>>(tm, p);

    return (int)cudaGetLastError();
}
