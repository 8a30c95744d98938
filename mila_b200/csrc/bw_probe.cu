// bw_probe.cu — read-only HBM access-pattern probes (development/measurement hook, not a product
// path).  Answers, on the box, "what does this DRAM system like at a 10-100 MB footprint?" so the
// decode kernel's load pattern is chosen from measurements, not folklore.  Every probe reads the
// same [rows, row_bytes] row-major byte matrix exactly once with G persistent CTAs of 8 warps
// and folds the data into a dummy checksum.
//
//   0  linear      : grid-stride LDG.128, each warp instruction covers 512 contiguous bytes
//   1  tile16x64   : the gemv_flat pattern — per warp step 16 rows x 64 B (LDG.128, 4 lanes/row), depth 4
//   2  tile16x64d8 : same, depth 8
//   3  bulk16x256  : per-warp cp.async.bulk ring, unit = 16 rows x 256 B, 3 units
//   4  bulk16x512  : per-warp cp.async.bulk ring, unit = 16 rows x 512 B, 2 units
//   5  bulk16x1024 : per-warp cp.async.bulk ring, unit = 16 rows x 1024 B, 2 units (4 warps)
//   6  tile8x128   : per warp step 8 rows x 128 B (LDG.128, 8 lanes/row), depth 4
//   7  row512      : per warp step 1 row x 512 B (LDG.128, 32 lanes/row), depth 8
#include "common.cuh"

namespace milab200 {
namespace {

__device__ __forceinline__ uint32_t fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{ .reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra WAIT_%=; }"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// ---- LDG tile patterns: ROWS rows x (LPR lanes x 16 B) per warp step ---------------------------
template <int LPR, int DEPTH>
__global__ void __launch_bounds__(256, 2)
probe_tile_kernel(const uint8_t* __restrict__ base, int64_t rows, int64_t row_bytes, uint32_t* out)
{
    constexpr int RPS = 32 / LPR;            // rows per load instruction
    constexpr int SPAN = LPR * 16;           // bytes per row per step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rr = lane / LPR, cc = lane % LPR;
    const int G = gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * rows / G, r_end = (int64_t)(blockIdx.x + 1) * rows / G;
    const int64_t tiles = (r_end - r_begin + 2 * RPS - 1) / (2 * RPS);      // two instructions per step like the GEMV
    const int64_t steps = row_bytes / SPAN;
    const int64_t total = tiles * steps;
    const int64_t a0 = warp * total / 8, a1 = (warp + 1) * total / 8;
    uint4 buf[DEPTH][2];
    uint32_t acc = 0;
    auto issue = [&](int64_t i, int slot) {
        const int64_t tile = i / steps, st = i - tile * steps;
        const int64_t r0 = r_begin + tile * 2 * RPS + rr, r1 = r0 + RPS;
        buf[slot][0] = make_uint4(0, 0, 0, 0); buf[slot][1] = make_uint4(0, 0, 0, 0);
        if (r0 < r_end) buf[slot][0] = ldg_stream_v4(base + r0 * row_bytes + st * SPAN + cc * 16);
        if (r1 < r_end) buf[slot][1] = ldg_stream_v4(base + r1 * row_bytes + st * SPAN + cc * 16);
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) if (a0 + d < a1) issue(a0 + d, d);
    for (int64_t b = a0; b < a1; b += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (b + d < a1) {
                acc ^= fold(buf[d][0]) ^ fold(buf[d][1]);
                if (b + d + DEPTH < a1) issue(b + d + DEPTH, d);
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(256, 2)
probe_linear_kernel(const uint8_t* __restrict__ base, int64_t bytes, uint32_t* out)
{
    const int64_t n = bytes / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint4 a = ldg_stream_v4(base + i * 16), b = ldg_stream_v4(base + (i + stride) * 16);
        const uint4 c = ldg_stream_v4(base + (i + 2 * stride) * 16), d = ldg_stream_v4(base + (i + 3 * stride) * 16);
        acc ^= fold(a) ^ fold(b) ^ fold(c) ^ fold(d);
    }
    for (; i < n; i += stride) acc ^= fold(ldg_stream_v4(base + i * 16));
    if (acc == 0x12345678u) out[0] = acc;
}

// ---- per-warp bulk-copy ring: unit = 16 rows x CH bytes ----------------------------------------
template <int CH, int UNITS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
probe_bulk_kernel(const uint8_t* __restrict__ base, int64_t rows, int64_t row_bytes, uint32_t* out)
{
    constexpr int PITCH = CH + 64;                       // bank-conflict-free row pitch for the consumer
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[WARPS * UNITS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* ring = smem + (size_t)warp * UNITS * 16 * PITCH;
    uint64_t* bar = bars + warp * UNITS;
    if (lane == 0) for (int u = 0; u < UNITS; ++u) mbar_init(&bar[u], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    const int G = gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * rows / G, r_end = (int64_t)(blockIdx.x + 1) * rows / G;
    const int64_t tiles = (r_end - r_begin + 15) / 16;
    const int64_t chunks = row_bytes / CH;
    const int64_t total = tiles * chunks;
    const int64_t a0 = warp * total / WARPS, a1 = (warp + 1) * total / WARPS;

    auto issue = [&](int64_t i, int u) {
        const int64_t tile = i / chunks, ch = i - tile * chunks;
        const int64_t r0 = r_begin + tile * 16;
        const int nrows = (int)min((int64_t)16, r_end - r0);
        if (lane == 0) mbar_expect_tx(&bar[u], (uint32_t)nrows * CH);
        __syncwarp();
        if (lane < nrows)
            bulk_g2s(ring + ((size_t)u * 16 + lane) * PITCH, base + (r0 + lane) * row_bytes + ch * CH, CH, &bar[u]);
    };
    for (int u = 0; u < UNITS; ++u) if (a0 + u < a1) issue(a0 + u, u);
    uint32_t acc = 0, phase = 0;
    int u = 0;
    const int g = lane >> 2, t = lane & 3;
    for (int64_t i = a0; i < a1; ++i) {
        mbar_wait(&bar[u], phase);
        // consume like the GEMV would: lane (g,t) reads rows g, g+8, 16 B at 64*s + 16*t
        const uint8_t* urow = ring + (size_t)u * 16 * PITCH;
#pragma unroll
        for (int s = 0; s < CH / 64; ++s) {
            acc ^= fold(*reinterpret_cast<const uint4*>(urow + (size_t)g * PITCH + s * 64 + t * 16));
            acc ^= fold(*reinterpret_cast<const uint4*>(urow + (size_t)(g + 8) * PITCH + s * 64 + t * 16));
        }
        __syncwarp();
        if (i + UNITS < a1) issue(i + UNITS, u);
        if (++u == UNITS) { u = 0; phase ^= 1; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

}  // namespace
}  // namespace milab200

using namespace milab200;

extern "C" int milab200_probe_bw(const void* base, int64_t rows, int64_t row_bytes, int pattern,
                                      void* out, milab200_stream_t stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint8_t* b = static_cast<const uint8_t*>(base);
    uint32_t* o = static_cast<uint32_t*>(out);
    switch (pattern) {
        case 0: probe_linear_kernel<<<sms * 2, 256, 0, stream>>>(b, rows * row_bytes, o); break;
        case 1: probe_tile_kernel<4, 4><<<sms * 2, 256, 0, stream>>>(b, rows, row_bytes, o); break;
        case 2: probe_tile_kernel<4, 8><<<sms * 2, 256, 0, stream>>>(b, rows, row_bytes, o); break;
        case 6: probe_tile_kernel<8, 4><<<sms * 2, 256, 0, stream>>>(b, rows, row_bytes, o); break;
        case 7: probe_tile_kernel<32, 8><<<sms * 2, 256, 0, stream>>>(b, rows, row_bytes, o); break;
        case 3: {
            constexpr int smem = 8 * 3 * 16 * (256 + 64);
            cudaFuncSetAttribute(probe_bulk_kernel<256, 3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe_bulk_kernel<256, 3, 8><<<sms, 256, smem, stream>>>(b, rows, row_bytes, o); break;
        }
        case 4: {
            constexpr int smem = 8 * 2 * 16 * (512 + 64);
            cudaFuncSetAttribute(probe_bulk_kernel<512, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe_bulk_kernel<512, 2, 8><<<sms, 256, smem, stream>>>(b, rows, row_bytes, o); break;
        }
        case 5: {
            constexpr int smem = 4 * 2 * 16 * (1024 + 64);
            cudaFuncSetAttribute(probe_bulk_kernel<1024, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe_bulk_kernel<1024, 2, 4><<<sms, 128, smem, stream>>>(b, rows, row_bytes, o); break;
        }
        default: return MILAB200_E_INVALID_ARGUMENT;
    }

    return (int)cudaGetLastError();
}
