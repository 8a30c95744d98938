// tp.cu — tensor-parallel plumbing of the decode path: the per-rank exchange buffer of the fused row-parallel
// all-reduce (decode_tc.cu epilogue) and the row-parallel entry points.
//
// The reference has no multi-GPU code at all (no NCCL/MPI symbol; tensor parallelism is a
// roadmap bullet, ROADMAP.md:278), so this is new surface.  One process per GPU; every rank allocates one
// exchange buffer with cudaMalloc, exports it with CUDA IPC, and maps every peer's buffer (NVLink P2P).  The
// host side (Python over torch.distributed, or any launcher) only has to all-gather the 64-byte handles.
//
// A row-parallel Linear (o_proj / down: the K dimension is sharded, FP4 groups never straddle shards;
// SURVEY.md §8e) computes FP32 partial rows on every rank; at M <= 16 the payload is 2*M*N*P bytes — tens of
// KB — so an NCCL all-reduce is pure latency (~10 us measured inside a graph at P = 2, longer than the GEMV
// itself).  Here the GEMV epilogue pushes the partials straight into the peers' buffers (8-byte {value, epoch}
// words, the "LL" wire format) and sums what the peers pushed: one kernel, no extra launch, one NVLink one-way
// latency, and with programmatic dependent launch the wait overlaps the next layer's weight prefetch.
#include <cstring>
#include <dlfcn.h>
#include <new>

#include "gemv_common.cuh"

namespace milab200 {
using namespace gemv;

int try_decode_tc(int fmt, __nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                  int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu);
int try_decode_mx4(__nv_bfloat16*, const __nv_bfloat16*, const uint8_t*, const float*, const __nv_bfloat16*,
                   int, int, int, cudaStream_t, int*, const TpExchange* tp, int glu);

int launch_gemv_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemv_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);
int launch_gemm_fp8(void*, const void*, const void*, const float*, const void*, int, int, int, cudaStream_t);
int launch_gemm_fp4(void*, const void*, const void*, const float*, const void*, int, int, int, int, cudaStream_t);

struct TpContext {
    int rank = 0, world = 1, nmax = 0;
    void* local = nullptr;
    size_t bytes = 0;
    void* opened[kTpMaxWorld] = {};
    bool connected = false;
    TpExchange view;
};

namespace {
size_t epoch_bytes(int nmax) { return (size_t)nmax * sizeof(uint32_t); }
size_t data_bytes(int world, int nmax) { return (size_t)2 * world * kMaxTok * nmax * sizeof(uint2); }

void fill_view(TpContext* c, int q, void* base)
{
    auto* b = static_cast<uint8_t*>(base);
    c->view.data[q] = reinterpret_cast<uint2*>(b + epoch_bytes(c->nmax));
    if (q == c->rank) c->view.row_epoch = reinterpret_cast<uint32_t*>(b);
}
}  // namespace

// decode_chain.cu: the peer-memory view of a connected context (null when not connected)
const TpExchange* tp_context_view(const void* ctx, int* nmax)
{
    auto* c = static_cast<const TpContext*>(ctx);
    if (!c || !c->connected) return nullptr;
    if (nmax) *nmax = c->nmax;
    return &c->view;
}

}  // namespace milab200

// ---- batched (or any-M) row-parallel Linear + NCCL all-reduce ---------------------------------------------------
// "opaque ncclComm_t + stream for the fused row-parallel + allreduce entry point".  The GEMM runs on this
// rank's K shard (bias on comm rank 0 only: it must be added once), then ncclAllReduce sums the BF16 partials in place
// on the same stream.  NCCL is not linked: its two entry points are resolved from the process image at first use (the
// host application — Mila, or torch — has NCCL loaded if it owns an ncclComm_t), so libmila_b200_linear.so carries no
// NCCL dependency for single-GPU users.
namespace {
typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclCommUserRankFn)(void*, int*);
NcclAllReduceFn g_nccl_allreduce = nullptr;
NcclCommUserRankFn g_nccl_rank = nullptr;
bool resolve_nccl()
{
    if (g_nccl_allreduce && g_nccl_rank) return true;
    void* h = nullptr;
    const char* names[] = { nullptr, "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        h = n ? dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD) : dlopen(nullptr, RTLD_NOW);
        if (!h) continue;
        auto a = reinterpret_cast<NcclAllReduceFn>(dlsym(h, "ncclAllReduce"));
        auto r = reinterpret_cast<NcclCommUserRankFn>(dlsym(h, "ncclCommUserRank"));
        if (a && r) { g_nccl_allreduce = a; g_nccl_rank = r; return true; }
    }
    return false;
}
constexpr int kNcclBfloat16 = 9, kNcclSum = 0;      // ncclDataType_t / ncclRedOp_t values (nccl.h, stable across 2.x)

}  // namespace


using namespace milab200;

extern "C" {

int milab200_tp_create(int rank, int world, int max_out_features, void** ctx_out)
{
    if (!ctx_out || world < 1 || world > kTpMaxWorld || rank < 0 || rank >= world || max_out_features <= 0)
        return MILAB200_E_INVALID_ARGUMENT;
    auto* c = new (std::nothrow) TpContext();
    if (!c) return MILAB200_E_INVALID_ARGUMENT;
    c->rank = rank; c->world = world; c->nmax = (max_out_features + 127) / 128 * 128;
    c->bytes = epoch_bytes(c->nmax) + data_bytes(world, c->nmax);
    cudaError_t e = cudaMalloc(&c->local, c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { if (c->local) cudaFree(c->local); delete c; cudaGetLastError(); return (int)e; }
    c->view.world = world; c->view.rank = rank; c->view.nmax = c->nmax;
    fill_view(c, rank, c->local);
    c->connected = (world == 1);
    *ctx_out = c;
    return 0;
}

int milab200_tp_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int milab200_tp_export(void* ctx, void* handle_out)
{
    auto* c = static_cast<TpContext*>(ctx);
    if (!c || !handle_out) return MILAB200_E_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    MILAB200_RETURN_IF_CUDA(cudaIpcGetMemHandle(&h, c->local));
    std::memcpy(handle_out, &h, sizeof(h));
    return 0;
}

/* all_handles: world consecutive handles in rank order (this rank's own entry is ignored) */
int milab200_tp_connect(void* ctx, const void* all_handles)
{
    auto* c = static_cast<TpContext*>(ctx);
    if (!c || !all_handles) return MILAB200_E_INVALID_ARGUMENT;
    if (c->connected) return 0;
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const uint8_t*>(all_handles) + (size_t)q * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        MILAB200_RETURN_IF_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->opened[q] = ptr;
        fill_view(c, q, ptr);
    }
    c->connected = true;
    return 0;
}

int milab200_tp_destroy(void* ctx)
{
    auto* c = static_cast<TpContext*>(ctx);
    if (!c) return 0;
    for (int q = 0; q < c->world; ++q)
        if (c->opened[q]) cudaIpcCloseMemHandle(c->opened[q]);
    if (c->local) cudaFree(c->local);
    delete c;
    return 0;
}

static int rowparallel(int fmt, void* out, const void* act, const void* w, const float* scales, const void* bias,
                       int M, int K, int N, void* ctx, milab200_stream_t stream)
{
    auto* c = static_cast<TpContext*>(ctx);
    if (!c || !out || !act || !w || !scales || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (!c->connected) return MILAB200_E_INVALID_ARGUMENT;
    if (M > kMaxTok || N > c->nmax) return MILAB200_E_BAD_SHAPE;       // batched row-parallel: GEMM + NCCL (caller)
    int status = 0;
    if (fmt == kFp4G128 &&
        try_decode_mx4(static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(act),
                       static_cast<const uint8_t*>(w), scales, static_cast<const __nv_bfloat16*>(bias), M, K, N,
                       static_cast<cudaStream_t>(stream), &status, &c->view, 0) == 0)
        return status;
    if (try_decode_tc(fmt, static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(act),
                      static_cast<const uint8_t*>(w), scales, static_cast<const __nv_bfloat16*>(bias), M, K, N,
                      static_cast<cudaStream_t>(stream), &status, &c->view, 0) != 0)
        return MILAB200_E_BAD_SHAPE;                                    // needs K % 128 == 0 and an sm_100 device
    return status;
}

int milab200_w8a16_gemm_rowparallel(void* out, const void* act, const void* w, const float* scales, const void* bias,
                                    int M, int K_local, int N, void* tp_ctx, milab200_stream_t stream)
{
    return rowparallel(kFp8, out, act, w, scales, bias, M, K_local, N, tp_ctx, stream);
}

int milab200_fp4a16_gemm_rowparallel(void* out, const void* act, const void* w, const float* scales, const void* bias,
                                     int M, int K_local, int N, int group_size, void* tp_ctx, milab200_stream_t stream)
{
    if (group_size != 128) return (group_size == 64) ? MILAB200_E_BAD_SHAPE : MILAB200_E_UNSUPPORTED_GROUP;
    return rowparallel(kFp4G128, out, act, w, scales, bias, M, K_local, N, tp_ctx, stream);
}

static int rowparallel_nccl(int fmt, void* out, const void* act, const void* w, const float* scales, const void* bias,
                            int M, int K, int N, int g, void* comm, milab200_stream_t stream)
{
    if (!out || !act || !w || !scales || !comm || M <= 0 || K <= 0 || N <= 0) return MILAB200_E_INVALID_ARGUMENT;
    if (!resolve_nccl()) return MILAB200_E_NO_NCCL;
    int rank = 0;
    if (g_nccl_rank(comm, &rank) != 0) return MILAB200_E_NO_NCCL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const void* b = (rank == 0) ? bias : nullptr;
    int rc;
    if (fmt == kFp8) rc = (M <= kMaxTok) ? launch_gemv_fp8(out, act, w, scales, b, M, K, N, st) : launch_gemm_fp8(out, act, w, scales, b, M, K, N, st);
    else             rc = (M <= kMaxTok) ? launch_gemv_fp4(out, act, w, scales, b, M, K, N, g, st) : launch_gemm_fp4(out, act, w, scales, b, M, K, N, g, st);
    if (rc != 0) return rc;
    const int nrc = g_nccl_allreduce(out, out, (size_t)M * (size_t)N, kNcclBfloat16, kNcclSum, comm, st);
    return nrc == 0 ? 0 : MILAB200_E_NO_NCCL;
}

int milab200_w8a16_gemm_rowparallel_nccl(void* out, const void* act, const void* w, const float* scales, const void* bias,
                                         int M, int K_local, int N, void* nccl_comm, milab200_stream_t stream)
{
    return rowparallel_nccl(kFp8, out, act, w, scales, bias, M, K_local, N, 0, nccl_comm, stream);
}

int milab200_fp4a16_gemm_rowparallel_nccl(void* out, const void* act, const void* w, const float* scales, const void* bias,
                                          int M, int K_local, int N, int group_size, void* nccl_comm, milab200_stream_t stream)
{
    if (group_size != 64 && group_size != 128) return MILAB200_E_UNSUPPORTED_GROUP;
    return rowparallel_nccl(group_size == 128 ? kFp4G128 : kFp4G64, out, act, w, scales, bias, M, K_local, N, group_size, nccl_comm, stream);
}

}  // extern "C"
