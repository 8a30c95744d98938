/*
 * mila_b200_linear.h — C-ABI of libmila_b200_linear.so
 *
 * B200-native (sm_100a) replacement for the kernel launchers behind Mila's quantized
 * Linear<TWeightQuant> forward path.  Each entry point replaces one free function of
 * namespace Mila::Dnn::Compute::Cuda::Linear (cited per function; paths relative to
 * Mila/Src/Dnn/Compute/Devices/Cuda/Operations/Linear/, "K/" = Kernels/).
 *
 * Conventions (same contract as the reference launchers):
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as an opaque pointer
 *   - everything is enqueued on `stream`; no entry point synchronises or allocates, all are
 *     CUDA-graph-capture safe (the *_host_* convenience entries at the bottom are the only
 *     exception and say so)
 *   - the callee owns nothing; buffers are caller-owned device allocations (>= 16-byte aligned)
 *   - return value: 0 on success, a cudaError_t (> 0) for CUDA failures, or one of the
 *     MILAB200_E_* codes (< 0) for argument errors.  Never throws.  The inline C++ wrappers in
 *     include/mila_b200/Kernels/ turn non-zero into the exception types the reference throws.
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef MILA_B200_LINEAR_H
#define MILA_B200_LINEAR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* milab200_stream_t;          /* cudaStream_t */

#define MILAB200_OK                    0
#define MILAB200_E_INVALID_ARGUMENT   -1   /* null pointer / non-positive size                      */
#define MILAB200_E_UNSUPPORTED_GROUP  -2   /* group_size not in {64,128}  (reference: runtime_error) */
#define MILAB200_E_BAD_SHAPE          -3   /* K % group_size != 0, K % 8 != 0 (reference: assert)    */
#define MILAB200_E_NO_DEVICE          -4   /* no CUDA device / wrong architecture (needs sm_100)     */
#define MILAB200_E_NO_NCCL            -5   /* *_rowparallel_nccl: NCCL not loaded in this process, or an NCCL call failed */

/* ABI version, bumped on any signature change. */
int         milab200_abi_version(void);
/* One-time per-device preparation (stream-K workspace of the decode kernel).  Optional: the first
 * forward does it lazily, but device memory cannot be allocated while a stream is being captured,
 * so call this (or run one eager forward) before capturing forwards into a CUDA graph.  Must not
 * be called during a capture.  Returns MILAB200_E_NO_DEVICE without an sm_100 device. */
int         milab200_init(void);
/* Sizes the library-owned activation workspace of the batched tensor-core path (two E4M3 planes of
 * the activations, 2*ceil128(max_tokens)*max_in_features + 4*ceil128(max_tokens) bytes) for the current
 * device — the counterpart of the reference context's grow-only scratch that CudaLinearOp carves its
 * FP8 activation buffer from (LIN/CudaLinearOp.ixx:660-676, CudaExecutionContext.ixx:164-270).
 * Optional: forwards grow it lazily, but not inside a stream capture, so call this before capturing
 * a forward with outer_size > 32.  Must not be called during a capture. */
int         milab200_reserve_prefill(int max_tokens, int max_in_features);
/* Human-readable message for a return code (static storage). */
const char* milab200_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Load-time quantizers
 * ------------------------------------------------------------------------------------------ */

/* Replaces cuda_quantize_fp8_per_channel — K/Quantization/CudaFp8WeightQuantization.cuh:56-63
 * (impl .cu:209-249, kernel :57-121).
 * Async H2D of the BF16 blob [N,K] into dev_staging (>= N*K*2 bytes), then per row:
 * absmax -> scale = absmax/448 (1 if 0) -> W8 = e4m3_satfinite_rn(w * (1/scale)).
 * src may be pinned or pageable host memory.  Outputs are bit-exact with the reference. */
int milab200_quantize_fp8_per_channel(const void* src_bf16_host, void* dst_fp8, float* dst_scales,
                                      int64_t out_features, int64_t in_features,
                                      void* dev_staging, milab200_stream_t stream);

/* Replaces cuda_quantize_fp4_per_group — K/Quantization/CudaFp4WeightQuantization.cuh:52-60
 * (impl .cu:184-223, kernel :83-144, encoder :54-70).  group_size in {64,128}.
 * packed [N,K/2] (low nibble = even column), scales [N,K/group_size].  Bit-exact. */
int milab200_quantize_fp4_per_group(const void* src_bf16_host, void* dst_packed, float* dst_scales,
                                    int64_t out_features, int64_t in_features, int group_size,
                                    void* dev_staging, milab200_stream_t stream);

/* Same quantizers with the BF16 source already resident on the device (no staging copy).
 * New surface (device-side quantize-on-load); same arithmetic. */
int milab200_quantize_fp8_per_channel_device(const void* src_bf16_dev, void* dst_fp8, float* dst_scales,
                                             int64_t out_features, int64_t in_features,
                                             milab200_stream_t stream);
int milab200_quantize_fp4_per_group_device(const void* src_bf16_dev, void* dst_packed, float* dst_scales,
                                           int64_t out_features, int64_t in_features, int group_size,
                                           milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Decode (memory-bound GEMV) — M = 1 in the reference; the batched entries below route
 * 2 <= M <= 16 to the same kernels.
 * ------------------------------------------------------------------------------------------ */

/* Replaces cuda_matvec_decode_bf16_qfp8 — K/Linear.cuh:49-57 (impl K/MatVec/CudaMatVecBias.Bf16.cu:527,
 * kernel :198).  y[OC] = bf16( scale[oc] * sum_c x[c]*f32(W8[oc,c]) + bias[oc] ).  C % 8 == 0. */
int milab200_matvec_decode_bf16_qfp8(void* y_bf16, const void* x_bf16, const void* weight_fp8,
                                     const float* scales, const void* bias_bf16,
                                     int C, int OC, milab200_stream_t stream);

/* Replaces cuda_matvec_decode_bf16_qfp4 — K/Linear.cuh:72-81 (impl Bf16.cu:545, kernels :271,:376).
 * y[OC] = bf16( sum_c x[c]*lut[nib]*scale[oc,c/g] + bias[oc] ).  C % group_size == 0. */
int milab200_matvec_decode_bf16_qfp4(void* y_bf16, const void* x_bf16, const void* weights_packed,
                                     const float* scales, const void* bias_bf16,
                                     int C, int OC, int group_size, milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Batched forward (any M >= 1).  M <= 16 runs the decode GEMV, larger M the tensor-core GEMM.
 * ------------------------------------------------------------------------------------------ */

/* Replaces cuda_w8a16_gemm — K/W8A16Gemm/CudaW8A16Gemm.cuh:59-68 (impl .cu:134, kernel :62).
 * out[M,N] = bf16( act[M,K] * (f32(W8[N,K]) * scale[n])^T + bias[n] ).
 * outer_size > 32 (and in_features % 128 == 0) runs the TMA + tcgen05 kernel of prefill_tc.cu, which
 * stages the activations in a library-owned per-device workspace: one batched forward at a time per
 * device (the reference is single-stream, CudaExecutionContext.ixx:368).
 * 9 <= outer_size <= 16 enqueues two kernels on `stream` (an activation pre-pass that splits the
 * activations once into the same kind of workspace — four rotating regions —, then the decode kernel);
 * both are capture-safe and the same single-stream rule applies. */
int milab200_w8a16_gemm(void* out_bf16, const void* act_bf16, const void* weight_fp8,
                        const float* scales, const void* bias_bf16,
                        int outer_size, int in_features, int out_features, milab200_stream_t stream);

/* Replaces cuda_fp4a16_gemm — K/W4A16Gemm/CudaW4A16Gemm.cuh:119-129 (impl .cu:367, kernel :88)
 * and cuda_fp4a16_gemm_wmma — K/W4A16Gemm/CudaW4A16Gemm.Wmma.cuh:46-56 (impl .Wmma.cu:297).
 * Unsupported group_size is an error here (the reference silently does nothing, .cu:361-363). */
int milab200_fp4a16_gemm(void* out_bf16, const void* act_bf16, const void* weights_packed,
                         const float* scales, const void* bias_bf16,
                         int outer_size, int in_features, int out_features, int group_size,
                         milab200_stream_t stream);
int milab200_fp4a16_gemm_wmma(void* out_bf16, const void* act_bf16, const void* weights_packed,
                              const float* scales, const void* bias_bf16,
                              int outer_size, int in_features, int out_features, int group_size,
                              milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Gate|up Linear with the gated activation fused into its epilogue (the step
 * right behind the path).  Mila runs fc_gate_up (ONE Linear, rows [0,H) = gate, [H,2H) = up:
 * Gemma.Block.ixx:347, Llama.Block.ixx:883) and then a separate GeGLU / SwiGLU kernel over its
 * [M,2H] BF16 output.  These entries produce the activation's [M,H] output directly; same
 * arithmetic as the two-kernel sequence (projections rounded to BF16 first, activation in FP32
 * with the reference's expressions).  Fused for M <= 16 (decode kernels: in_features % 128 == 0, FP8 or FP4
 * g=128, H/128 >= ~3/4 of the SM count) and for M > 32 with FP8 weights (batched kernel: whole-K tiles,
 * exact activation planes, CTA pairs); otherwise the Linear is written to gate_up_scratch [M,2H]
 * (the Linear's own output tensor in the reference; may be NULL only when the fused path applies)
 * and the stand-alone activation kernel below follows.
 * ------------------------------------------------------------------------------------------ */
#define MILAB200_GLU_GEGLU_TANH 1     /* Activations/Geglu/Kernels/Geglu.cu:42-61  (Gemma)  */
#define MILAB200_GLU_SWIGLU     2     /* Activations/Swiglu/Kernels/Swiglu.Bf16.cu:135-230 (Llama) */
int milab200_w8a16_gemm_glu(void* out_bf16, void* gate_up_scratch_bf16, const void* act_bf16,
                            const void* weight_fp8, const float* scales, const void* bias_bf16,
                            int outer_size, int in_features, int hidden /* = out_features / 2 */,
                            int glu_kind, milab200_stream_t stream);
int milab200_fp4a16_gemm_glu(void* out_bf16, void* gate_up_scratch_bf16, const void* act_bf16,
                             const void* weights_packed, const float* scales, const void* bias_bf16,
                             int outer_size, int in_features, int hidden, int group_size,
                             int glu_kind, milab200_stream_t stream);
/* Replace cuda_geglu_forward_bf16 — Activations/Geglu/Kernels/Geglu.cuh:29-32 (impl .cu:81-96) and
 * cuda_swiglu_forward_bf16 — Activations/Swiglu/Kernels/Swiglu.cuh:30-34 (impl Swiglu.Bf16.cu:135-230).
 * Y[N] with N = tokens * half_width; X [tokens, 2*half_width].  Bit-exact with the reference kernels. */
int milab200_geglu_forward_bf16(void* Y_bf16, const void* X_bf16, int N, int half_width, milab200_stream_t stream);
int milab200_swiglu_forward_bf16(void* Y_bf16, const void* X_bf16, int N, int half_width, milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FP8 tied-table token embedding: the gather that shares the lm_head's
 * PerChannelFp8 table and scales.  Replace cuda_token_embedding_forward_bf16_qfp8 /
 * cuda_token_embedding_decode_bf16_qfp8 — Embeddings/Kernels/TokenEmbedding.cuh:46-52 (impl
 * TokenEmbedding.Fp8.cu:34-147).  Y[bt,:] = bf16( f32(wte_fp8[X[bt],:]) * scales[X[bt]] ); C % 8 == 0.
 * Bit-exact.  The table is built with milab200_quantize_fp8_per_channel called per row chunk, as
 * CudaTokenEmbeddingOp.Quantize.ixx:63-113 does with the reference quantizer.
 * ------------------------------------------------------------------------------------------ */
int milab200_token_embedding_forward_bf16_qfp8(void* Y_bf16, const int* X_ids, const void* wte_fp8,
                                               const float* scales, int B, int T, int C, milab200_stream_t stream);
int milab200_token_embedding_decode_bf16_qfp8(void* Y_bf16, const int* X_ids, const void* wte_fp8,
                                              const float* scales, int B, int C, milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tensor parallelism (new surface: the reference is single-GPU; tensor parallelism is a roadmap
 * bullet, ROADMAP.md:278).  One process per GPU.  Column-parallel shards (QKV / gate / up: row
 * slices of the weight) need nothing new — call the entries above on the shard.  A row-parallel
 * shard (o_proj / down: K slice holding whole FP4 groups, FP8 row scales from the unsharded
 * quantisation) produces partial rows that must be summed over the ranks; for decode (M <= 16)
 * the *_rowparallel entries do that inside the GEMV epilogue over NVLink peer memory (one-shot
 * all-reduce, no extra launch).  For M > 16 they return MILAB200_E_BAD_SHAPE: run the batched
 * entry on the shard and all-reduce the BF16 partials with NCCL (bandwidth-bound regime).
 *
 * Setup: every rank creates a context (allocates its exchange buffer), exports its 64-byte CUDA
 * IPC handle, the host all-gathers the handles (any transport), every rank connects.
 * ------------------------------------------------------------------------------------------ */
int milab200_tp_create(int rank, int world_size, int max_out_features, void** ctx_out);
int milab200_tp_handle_bytes(void);                              /* sizeof(cudaIpcMemHandle_t) = 64 */
int milab200_tp_export(void* ctx, void* handle_out);
int milab200_tp_connect(void* ctx, const void* all_handles);     /* world_size handles, rank order */
int milab200_tp_destroy(void* ctx);
/* out[M,N] (identical on every rank) = bf16( sum_ranks act_r[M,K_local] * dequant(W_r[N,K_local])^T + bias ).
 * Every rank of the group must enqueue the same sequence of *_rowparallel calls. */
int milab200_w8a16_gemm_rowparallel(void* out_bf16, const void* act_bf16, const void* weight_fp8_shard,
                                    const float* scales, const void* bias_bf16,
                                    int outer_size, int in_features_local, int out_features,
                                    void* tp_ctx, milab200_stream_t stream);
int milab200_fp4a16_gemm_rowparallel(void* out_bf16, const void* act_bf16, const void* weights_packed_shard,
                                     const float* scales_shard, const void* bias_bf16,
                                     int outer_size, int in_features_local, int out_features, int group_size,
                                     void* tp_ctx, milab200_stream_t stream);
/* Any outer_size (the batched regime above all): the Linear on this rank's K shard — bias on communicator rank 0 only —
 * followed by ncclAllReduce(sum, BF16, in place) on `stream`.  `nccl_comm` is the caller's ncclComm_t; NCCL itself is
 * resolved from the process image at first use (no link-time dependency), MILAB200_E_NO_NCCL if it is not loaded. */
int milab200_w8a16_gemm_rowparallel_nccl(void* out_bf16, const void* act_bf16, const void* weight_fp8_shard,
                                         const float* scales, const void* bias_bf16,
                                         int outer_size, int in_features_local, int out_features,
                                         void* nccl_comm, milab200_stream_t stream);
int milab200_fp4a16_gemm_rowparallel_nccl(void* out_bf16, const void* act_bf16, const void* weights_packed_shard,
                                          const float* scales_shard, const void* bias_bf16,
                                          int outer_size, int in_features_local, int out_features, int group_size,
                                          void* nccl_comm, milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * RMSNorm -> Linear (second half).  Mila normalises into a BF16 tensor and the Linear reads it
 * back (Gemma.Block.ixx:209-210,347-349).  milab200_rmsnorm_forward_bf16 replaces cuda_rmsnorm_forward_bf16 —
 * Normalizations/RmsNorm/Kernels/RmsNorm.cuh:125 (kernel RmsNorm.Bf16.cu:19-73) — bit for bit (weight_offset = 1 for
 * Gemma's (1 + weight), Gemma.Block.ixx:18-20).  milab200_rmsnorm_{w8a16,fp4a16}_gemm compute
 *     out[M,N] = Linear( RMSNorm(act[M,K]; norm_weight, norm_bias, epsilon, weight_offset) )
 * with the normalisation done inside the Linear's activation path (same reduction order, same expression, same BF16
 * rounding as the stand-alone kernel), so the result equals the two-kernel sequence bit for bit and the normalised
 * tensor is never written.  normed_scratch [M,K] BF16 is only touched for shapes the fused routes do not take
 * (in_features % 128 != 0, group 64) — may be NULL if the caller knows better.
 * ------------------------------------------------------------------------------------------ */
int milab200_rmsnorm_forward_bf16(void* Y_bf16, void* rstd_bf16_or_null, const void* X_bf16, const void* weight_bf16,
                                  const void* bias_bf16, int outer_size, int inner_size, int norm_dim,
                                  float epsilon, float weight_offset, milab200_stream_t stream);
int milab200_rmsnorm_w8a16_gemm(void* out_bf16, void* normed_scratch, const void* act_bf16, const void* norm_weight_bf16,
                                const void* norm_bias_bf16, float epsilon, float weight_offset,
                                const void* weight_fp8, const float* scales, const void* bias_bf16,
                                int outer_size, int in_features, int out_features, milab200_stream_t stream);
int milab200_rmsnorm_fp4a16_gemm(void* out_bf16, void* normed_scratch, const void* act_bf16, const void* norm_weight_bf16,
                                 const void* norm_bias_bf16, float epsilon, float weight_offset,
                                 const void* weights_packed, const float* scales, const void* bias_bf16,
                                 int outer_size, int in_features, int out_features, int group_size, milab200_stream_t stream);

/* RMSNorm -> gate|up Linear -> GeGLU / SwiGLU: the front half of Mila's MLP block (ln_2 -> fc_gate_up -> activation,
 * Gemma.Block.ixx:209-210,347-349; Llama.Block.ixx:883) as ONE call —
 *     out[M,H] = GLU( Linear_{2H x K}( RMSNorm(act[M,K]) ) ),  weight rows [0,H) gate, [H,2H) up
 * On the decode routes (M <= 16, in_features % 128 == 0, FP8 or FP4 g = 128) it is one launch: the norm runs in the
 * activation converters, the gated activation in the epilogue; result = the reference's three-kernel sequence
 * (cuda_rmsnorm_forward_bf16 -> Linear -> cuda_geglu/swiglu_forward_bf16) bit for bit.  Other shapes run that sequence
 * through normed_scratch [M,K] and gate_up_scratch [M,2H] (BF16; may be NULL only if the caller knows the route). */
int milab200_rmsnorm_w8a16_gemm_glu(void* out_bf16, void* gate_up_scratch, void* normed_scratch, const void* act_bf16,
                                    const void* norm_weight_bf16, const void* norm_bias_bf16, float epsilon, float weight_offset,
                                    const void* weight_fp8, const float* scales, const void* bias_bf16,
                                    int outer_size, int in_features, int hidden_features, int glu_kind, milab200_stream_t stream);
int milab200_rmsnorm_fp4a16_gemm_glu(void* out_bf16, void* gate_up_scratch, void* normed_scratch, const void* act_bf16,
                                     const void* norm_weight_bf16, const void* norm_bias_bf16, float epsilon, float weight_offset,
                                     const void* weights_packed, const float* scales, const void* bias_bf16,
                                     int outer_size, int in_features, int hidden_features, int group_size, int glu_kind,
                                     milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * PerGroupInt4<g> (GPTQ-style W4A16).  Replaces cuda_w4a16_gemm —
 * K/W4A16Gemm/CudaW4A16Gemm.cuh:73 (impl .cu:329, kernel :88-197; call sites LIN/CudaLinearOp.ixx:560,786,866).
 * weights_packed [N, K/2]: two unsigned INT4 per byte, low nibble = even k; scales [N, K/g] FP32;
 * zero_points [N, K/(2g)]: two INT4 per byte, low nibble = even group, or NULL (symmetric, zero = 8);
 * out[m,n] = bf16( sum_k act[m,k] * (nibble - zero) * scale + bias[n] ).  g in {64, 128}; any outer_size.
 * The policy has no quantizer in the reference (CudaLinearOp.ixx:385-391): checkpoints arrive pre-quantized.
 * ------------------------------------------------------------------------------------------ */
int milab200_w4a16_gemm(void* out_bf16, const void* act_bf16, const void* weights_packed, const float* scales,
                        const void* zero_points, const void* bias_bf16,
                        int outer_size, int in_features, int out_features, int group_size, milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Chained decode: a list of dependent decode Linears (M <= 16) as ONE persistent launch.
 *
 * New surface (the reference launches one kernel per Linear::forward and lists CUDA-graph decode as its next
 * lever, CHANGELOG.md:252-258).  On B200 a decode Linear streams its weights in 1-9 us and every kernel boundary
 * between two dependent Linears idles HBM for ~3 us; the chained kernel walks the list with device-side
 * dependency counters, so the weight stream of entry i + 1 runs under the epilogue of entry i.  Each entry is exactly
 * milab200_w8a16_gemm / milab200_fp4a16_gemm [+ _glu / _rowparallel] of the same arguments.
 *
 * Semantics: entries execute as if launched one after the other on `stream` (entry i reads its activations after
 * entry depends_on has completely finished; use i - 1 for plain stream order, -1 for "ready at launch", or an earlier
 * index when the caller knows the entries in between neither produce this entry's input nor are still reading a buffer
 * this entry writes).  One policy per chain (all PerChannelFp8 or all PerGroupFp4<128>), one outer_size, in_features %
 * 128 == 0.  create allocates (not capturable); forward = one memset + one kernel (capturable, replayable).
 * ------------------------------------------------------------------------------------------ */
typedef struct milab200_chain_linear {
    void*        out_bf16;        /* [M, out_features]  ([M, out_features / 2] when glu != 0) */
    const void*  act_bf16;        /* [M, in_features] */
    const void*  weight;          /* FP8 [N,K] or packed FP4 [N,K/2] */
    const float* scales;          /* [N] or [N, K/128] */
    const void*  bias_bf16;       /* [N] or NULL */
    int in_features, out_features;
    int group_size;               /* 0 = PerChannelFp8, 128 = PerGroupFp4<128> */
    int glu;                      /* 0, MILAB200_GLU_GEGLU_TANH or MILAB200_GLU_SWIGLU: weight is gate|up, rows [0,H) gate */
    int depends_on;               /* see above */
    void* tp_ctx;                 /* row-parallel shard: sum over the ranks in the epilogue (NULL = none) */
} milab200_chain_linear;
int milab200_chain_create(const milab200_chain_linear* linears, int count, int outer_size, void** chain_out);
int milab200_chain_forward(void* chain, milab200_stream_t stream);
int milab200_chain_destroy(void* chain);
/* tile height, k-splits (1 or 2) and row tiles the balanced decomposition picked for entry `index` */
int milab200_chain_describe(void* chain, int index, int* tile_rows, int* ksplits, int* tiles);
/* bring-up: per-layer role timestamps (count * 8 * grid int64, device memory) written by the next forwards; NULL = off */
int milab200_chain_set_timeline(void* chain, void* device_buf, int* grid_out);

/* ------------------------------------------------------------------------------------------
 * Staging / W4A8 helpers the reference's 2-phase paths call (kept so CudaLinearOp.ixx links
 * unchanged whichever toggles are set).  Bit-exact with the reference kernels.
 * ------------------------------------------------------------------------------------------ */

/* cuda_fp8_dequantize_to_bf16 — K/Fp8Prefill/CudaFp8Prefill.cuh (impl .cu:86, kernel :64). */
int milab200_fp8_dequantize_to_bf16(void* out_bf16, const void* weight_fp8, const float* scales,
                                    int out_features, int in_features, milab200_stream_t stream);
/* cuda_fp4_dequantize_to_bf16 — K/W4A16Gemm/CudaW4A16Gemm.cuh (impl .cu:403, kernel :210). */
int milab200_fp4_dequantize_to_bf16(void* out_bf16, const void* weights_packed, const float* scales,
                                    int out_features, int in_features, int group_size,
                                    milab200_stream_t stream);
/* cuda_compute_fp8_weight_scale — CudaW4A16Gemm.cuh (impl .cu:433, kernels :244,:285). */
int milab200_compute_fp8_weight_scale(float* weight_fp8_scale_out, const float* fp4_group_scales,
                                      int64_t num_scales, milab200_stream_t stream);
/* cuda_fp4_dequantize_to_fp8 — CudaW4A16Gemm.cuh (impl .cu:451, kernel :300). */
int milab200_fp4_dequantize_to_fp8(void* out_fp8, const void* weights_packed, const float* scales,
                                   const float* weight_fp8_scale, int out_features, int in_features,
                                   int group_size, milab200_stream_t stream);
/* cuda_quantize_bf16_to_fp8_per_token — K/Fp8Prefill/CudaFp8Prefill.cuh (impl .cu:165, kernel :116). */
int milab200_quantize_bf16_to_fp8_per_token(void* fp8_out, float* scales_out, const void* input_bf16,
                                            int outer_size, int in_features, milab200_stream_t stream);
/* cuda_fp8_apply_per_token_scales — CudaFp8Prefill.cuh (impl .cu:213, kernel :191). */
int milab200_fp8_apply_per_token_scales(void* output_bf16, const float* scales, const void* bias_bf16,
                                        int outer_size, int out_features, milab200_stream_t stream);
/* cuda_add_bias (BF16 overload) — CudaFp8Prefill.cuh (impl .cu:258, kernel :239). */
int milab200_add_bias_bf16(void* output_bf16, const void* bias_bf16,
                           int outer_size, int out_features, milab200_stream_t stream);
/* cuda_add_bias (FP32 overload) — CudaFp8Prefill.cuh:151. */
int milab200_add_bias_f32(float* output, const float* bias, int outer_size, int out_features, milab200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Runtime options and stream-order contract
 * ------------------------------------------------------------------------------------------ */

/* Route selection, the programmatic form of the MILAB200_* environment switches.  Every route
 * computes the same function and is parity-tested; the defaults are the measured best.  Names:
 *   "decode_tc" (1)          0: mma.sync decode kernels only
 *   "decode_mx4_max_m" (2)   largest M of the packed-nibble kind::mxf4 FP4 decode kernel (0 = off)
 *   "decode_streamk" (-1)    stream-K work decomposition: -1 auto, 0 off, 1 on
 *   "decode_presplit" (1)    M = 9..16: 1 activation pre-pass kernel, 0 in-kernel converter warps
 *   "decode_generic" (0)     1: every decode call takes the one-warp-per-row FP32 kernel (cross-check route)
 *   "prefill_tc" (1)         0: token-blocked decode kernels for M > 16
 *   "prefill_cta_group" (2)  2 CTA pairs (tcgen05 cta_group::2), 1 single-CTA tiles
 *   "prefill_act_planes" (2) batched path: 2 = activations split exactly into two E4M3 planes (conforming: 1e-2 of the
 *                            FP32 reference), 1 = ONE per-token-scaled E4M3 plane, the reference's own W4A8 activation
 *                            format (CudaFp8Prefill.cu:116-165, LIN/CudaLinearOp.ixx:660-714; its gate: 1e-1 of the row
 *                            maximum, Linear.Cuda.cpp:773) at twice the useful tensor rate.  Lossy, opt-in, outer_size
 *                            >= 256; FP8 weights, and FP4 g = 128 weights with out_features >= 256.
 *   "prefill_fp4_sum" (1)    batched FP4 path: both exact activation planes accumulate into the SAME accumulator columns
 *                            (lo = rn(v - hi); 256-token tiles) — 1 where those tiles fill the GPU without a k split,
 *                            2 whenever outer_size >= 256, 0 never (hi | lo columns, combined in the epilogue)
 *   "prefill_glu" (1)        gate|up Linear + GLU entries at outer_size > 32, FP8 weights: 1 = the activation in the epilogue of
 *                            the batched kernel (CTA pairs: gate rows | up rows, each CTA finishes half of the tokens, the
 *                            halves cross over distributed shared memory; bit-identical to the Linear followed by the
 *                            activation kernel), 0 = that sequence
 *   "rmsnorm_fast_reduction" (0) fused RMSNorm -> Linear entries: 1 = sum of squares in tree order with 128-bit loads
 *                            (~0.5 us instead of ~3 us on the dependency chain of a decode Linear); rstd then differs
 *                            from the reference's lane-strided order in the last FP32 bits and the fused result is no
 *                            longer bit-identical to the kernel sequence (still far inside 1e-2).  Opt-in.
 * Unknown name: MILAB200_E_INVALID_ARGUMENT. */
int milab200_set_option(const char* name, int value);

/* Stream order.  Every entry point is ordered on `stream` like the reference launchers, with ONE documented
 * relaxation: a decode launch may begin reading its WEIGHT bytes before the previous kernel on the stream has
 * finished (programmatic dependent launch; activations, scales-as-written-by-quantizers and outputs keep plain
 * order).  The library's own quantizers and weight-writing helpers disable that for the launch that follows them.
 * A caller that writes weight storage with kernels of its own (device-side copy, tied-table install) calls this
 * once after enqueuing them; the next decode launch on the current device then takes plain stream order.
 * MILAB200_PDL=0 in the environment disables the relaxation altogether. */
int milab200_note_weights_written(void);

/* ------------------------------------------------------------------------------------------
 * Introspection (used by bench.py's gpu_launches count and by the tests)
 * ------------------------------------------------------------------------------------------ */

/* Number of kernels this library has launched since load / since the last reset. */
uint64_t milab200_launch_count(void);
void     milab200_reset_launch_count(void);
/* Name of the kernel variant the last forward entry dispatched to (static storage). */
const char* milab200_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* MILA_B200_LINEAR_H */
