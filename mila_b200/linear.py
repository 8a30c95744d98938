"""Host-side mirror of Mila's quantized Linear component, driving the C-ABI.

Mirrors, for the quantized CUDA/BF16 instantiations only, the interface of
  Linear<DeviceType::Cuda, BF16, TWeightQuant>   Mila/Src/Dnn/Components/Linear/Linear.ixx:82-1092
  LinearConfig                                   Mila/Src/Dnn/Components/Linear/LinearConfig.ixx:39-204
  PerChannelFp8<> / PerGroupFp4<g>               Mila/Src/Dnn/Quantization/Weight/Policies.ixx:46-53,104-113
  CudaLinearOp<BF16,TWeightQuant>::forward       .../Operations/Linear/CudaLinearOp.ixx:535-880 (routing by M)
with the same method names, argument meaning and error behaviour, so that the parity tests read
like Mila/Tests/Dnn/Components/Linear/Linear.Cuda.cpp.  Exceptions map as
std::invalid_argument -> InvalidArgument(ValueError), std::logic_error -> LogicError,
std::runtime_error -> MilaB200Error(RuntimeError).

PyTorch is used for device memory, streams and dtype views only; every computation goes through
libmila_b200_linear.so.  No CPU fallback: constructing a Linear without CUDA raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import InvalidArgument, LogicError, MilaB200Error

__all__ = ["PerChannelFp8", "PerGroupFp4", "PerGroupInt4", "w4a16_forward", "LinearConfig", "TensorBlob", "Linear",
           "quantize_fp8_per_channel", "quantize_fp4_per_group", "linear_forward",
           "InvalidArgument", "LogicError", "MilaB200Error"]


# ---- weight-quantisation policies (Policies.ixx) -----------------------------------------

@dataclass(frozen=True)
class PerChannelFp8:
    """Policies.ixx:46-53 — FP8_E4M3 storage, one FP32 scale per output channel."""
    kIsQuantized: bool = True
    kStorageDtype: str = "FP8_E4M3"
    kScaleDtype: str = "FP32"
    kPerChannel: bool = True
    tag: str = "per_channel_fp8_e4m3"        # Core/LanguageModelConfig.ixx:104-114


@dataclass(frozen=True)
class PerGroupFp4:
    """Policies.ixx:104-113 — packed E2M1 nibbles in UINT8, one FP32 scale per group."""
    kQuantizationGroupSize: int = 128
    kIsQuantized: bool = True
    kStorageDtype: str = "UINT8"
    kScaleDtype: str = "FP32"
    kPerChannel: bool = False
    kIsFp4E2M1: bool = True

    @property
    def tag(self) -> str:
        return f"per_group_fp4_{self.kQuantizationGroupSize}"


@dataclass(frozen=True)
class PerGroupInt4:
    """Policies.ixx:71-79 — packed unsigned INT4 in UINT8 (low nibble = even column), one FP32 scale and one INT4 zero
    point per group (zero points optional: symmetric, zero = 8).  Pre-quantized checkpoints only: the reference has no
    quantizer for this policy (CudaLinearOp.ixx:385-391)."""
    kQuantizationGroupSize: int = 128
    kIsQuantized: bool = True
    kStorageDtype: str = "UINT8"
    kScaleDtype: str = "FP32"
    kPerChannel: bool = False
    kIsFp4E2M1: bool = False

    @property
    def tag(self) -> str:
        return f"per_group_int4_{self.kQuantizationGroupSize}"


class LinearConfig:
    """LinearConfig.ixx:39-204 (the fields the quantized forward path reads)."""

    def __init__(self, in_features: int, out_features: int):
        if in_features <= 0 or out_features <= 0:
            raise InvalidArgument("LinearConfig: features must be positive")
        self._in, self._out, self._bias = int(in_features), int(out_features), True

    def withBias(self, has_bias: bool) -> "LinearConfig":
        self._bias = bool(has_bias)
        return self

    def getInputFeatures(self) -> int: return self._in
    def getOutputFeatures(self) -> int: return self._out
    def hasBias(self) -> bool: return self._bias


@dataclass
class TensorBlob:
    """ITensorBlob (Tensors/Tensor.Serialization.ixx:34-): dtype tag + shape + host bytes.
    `data` is a CPU torch tensor (any dtype view of the raw bytes; pinned or pageable)."""
    dtype: str                 # "BF16" | "FP8_E4M3" | "UINT8" | "FP32"
    shape: tuple
    data: torch.Tensor


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _p(t: torch.Tensor | None):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


# ---- free functions: thin typed wrappers over the C-ABI ------------------------------------

def quantize_fp8_per_channel(w_bf16_host: torch.Tensor, device="cuda:0", staging: torch.Tensor | None = None):
    """cuda_quantize_fp8_per_channel (CudaFp8WeightQuantization.cuh:56-63).
    w_bf16_host: CPU bf16 [N,K].  Returns (W8 uint8 [N,K], scales f32 [N]) on `device`.
    A CUDA source tensor takes the *_device entry (no staging copy)."""
    N, K = w_bf16_host.shape
    dev = w_bf16_host.device if w_bf16_host.is_cuda else torch.device(device)
    q = torch.empty((N, K), dtype=torch.uint8, device=dev)
    s = torch.empty((N,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        if w_bf16_host.is_cuda:
            rc = L.milab200_quantize_fp8_per_channel_device(_p(w_bf16_host.contiguous()), _p(q), _p(s), N, K,
                                                            _stream_ptr(dev))
        else:
            w = w_bf16_host.contiguous()
            if staging is None:
                staging = torch.empty((N, K), dtype=torch.bfloat16, device=dev)
            rc = L.milab200_quantize_fp8_per_channel(_p(w), _p(q), _p(s), N, K, _p(staging), _stream_ptr(dev))
    _lib.check(rc, "quantize_fp8_per_channel")
    return q, s


def quantize_fp4_per_group(w_bf16_host: torch.Tensor, group_size: int = 128, device="cuda:0",
                           staging: torch.Tensor | None = None):
    """cuda_quantize_fp4_per_group (CudaFp4WeightQuantization.cuh:52-60).
    Returns (packed uint8 [N,K/2], scales f32 [N,K/g])."""
    N, K = w_bf16_host.shape
    dev = w_bf16_host.device if w_bf16_host.is_cuda else torch.device(device)
    q = torch.empty((N, K // 2), dtype=torch.uint8, device=dev)
    s = torch.empty((N, max(K // max(group_size, 1), 1)), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        if w_bf16_host.is_cuda:
            rc = L.milab200_quantize_fp4_per_group_device(_p(w_bf16_host.contiguous()), _p(q), _p(s), N, K,
                                                          group_size, _stream_ptr(dev))
        else:
            w = w_bf16_host.contiguous()
            if staging is None:
                staging = torch.empty((N, K), dtype=torch.bfloat16, device=dev)
            rc = L.milab200_quantize_fp4_per_group(_p(w), _p(q), _p(s), N, K, group_size, _p(staging),
                                                   _stream_ptr(dev))
    _lib.check(rc, "quantize_fp4_per_group")
    return q, s


def linear_forward(x: torch.Tensor, weight: torch.Tensor, scales: torch.Tensor, policy,
                   bias: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """CudaLinearOp::forward routing (CudaLinearOp.ixx:535-880) over our entry points:
    M == 1 -> cuda_matvec_decode_bf16_qfp{8,4}; M > 1 -> cuda_w8a16_gemm / cuda_fp4a16_gemm
    (kUseW8A16Gemm = kUseFusedFp4Gemm = true)."""
    K = x.shape[-1]
    M = x.numel() // K
    N = weight.shape[0]
    if out is None:
        out = torch.empty((*x.shape[:-1], N), dtype=torch.bfloat16, device=x.device)
    L = _lib.lib()
    st = _stream_ptr(x.device)
    with torch.cuda.device(x.device):
        if isinstance(policy, PerChannelFp8):
            if M == 1:
                rc = L.milab200_matvec_decode_bf16_qfp8(_p(out), _p(x), _p(weight), _p(scales), _p(bias), K, N, st)
            else:
                rc = L.milab200_w8a16_gemm(_p(out), _p(x), _p(weight), _p(scales), _p(bias), M, K, N, st)
        else:
            g = policy.kQuantizationGroupSize
            if M == 1:
                rc = L.milab200_matvec_decode_bf16_qfp4(_p(out), _p(x), _p(weight), _p(scales), _p(bias), K, N, g, st)
            else:
                rc = L.milab200_fp4a16_gemm(_p(out), _p(x), _p(weight), _p(scales), _p(bias), M, K, N, g, st)
    _lib.check(rc, "linear_forward")
    return out


def w4a16_forward(x: torch.Tensor, weight: torch.Tensor, scales: torch.Tensor, zero_points: torch.Tensor | None,
                  group_size: int = 128, bias: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """cuda_w4a16_gemm (CudaW4A16Gemm.cuh:73) — the PerGroupInt4 branch of CudaLinearOp::forward (CudaLinearOp.ixx:560,786,866)."""
    K = x.shape[-1]
    M = x.numel() // K
    N = weight.shape[0]
    if out is None:
        out = torch.empty((*x.shape[:-1], N), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().milab200_w4a16_gemm(_p(out), _p(x), _p(weight), _p(scales), _p(zero_points), _p(bias), M, K, N,
                                            group_size, _stream_ptr(x.device))
    _lib.check(rc, "w4a16_forward")
    return out


def rmsnorm_forward(x: torch.Tensor, weight: torch.Tensor | None, bias: torch.Tensor | None = None, eps: float = 1e-6,
                    weight_offset: float = 0.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """cuda_rmsnorm_forward_bf16 over the last dimension (RmsNorm.cuh:125; weight_offset = 1 for Gemma's (1 + w))."""
    K = x.shape[-1]
    M = x.numel() // K
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().milab200_rmsnorm_forward_bf16(_p(out), None, _p(x), _p(weight), _p(bias), M, 1, K, eps, weight_offset,
                                                      _stream_ptr(x.device))
    _lib.check(rc, "rmsnorm_forward")
    return out


def rmsnorm_linear_forward(x: torch.Tensor, norm_weight: torch.Tensor | None, norm_bias: torch.Tensor | None, eps: float,
                           weight_offset: float, weight: torch.Tensor, scales: torch.Tensor, policy,
                           bias: torch.Tensor | None = None, out: torch.Tensor | None = None,
                           scratch: torch.Tensor | None = None) -> torch.Tensor:
    """norm->forward followed by Linear->forward (Gemma.Block.ixx:209-210,347-349) as one call; `scratch` [M,K] is the
    norm's own output tensor, used only when the fused routes do not take the shape."""
    K = x.shape[-1]
    M = x.numel() // K
    N = weight.shape[0]
    if out is None:
        out = torch.empty((*x.shape[:-1], N), dtype=torch.bfloat16, device=x.device)
    if scratch is None:
        scratch = torch.empty((M, K), dtype=torch.bfloat16, device=x.device)
    L = _lib.lib()
    st = _stream_ptr(x.device)
    with torch.cuda.device(x.device):
        if isinstance(policy, PerChannelFp8):
            rc = L.milab200_rmsnorm_w8a16_gemm(_p(out), _p(scratch), _p(x), _p(norm_weight), _p(norm_bias), eps, weight_offset,
                                               _p(weight), _p(scales), _p(bias), M, K, N, st)
        else:
            rc = L.milab200_rmsnorm_fp4a16_gemm(_p(out), _p(scratch), _p(x), _p(norm_weight), _p(norm_bias), eps, weight_offset,
                                                _p(weight), _p(scales), _p(bias), M, K, N, policy.kQuantizationGroupSize, st)
    _lib.check(rc, "rmsnorm_linear_forward")
    return out


GLU_GEGLU_TANH, GLU_SWIGLU = 1, 2


def glu_forward(x: torch.Tensor, kind: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """cuda_geglu_forward_bf16 / cuda_swiglu_forward_bf16 (Activations/{Geglu,Swiglu}/Kernels): x [..., 2H] with the
    gate half first -> [..., H]."""
    H = x.shape[-1] // 2
    if out is None:
        out = torch.empty((*x.shape[:-1], H), dtype=torch.bfloat16, device=x.device)
    L = _lib.lib()
    fn = L.milab200_geglu_forward_bf16 if kind == GLU_GEGLU_TANH else L.milab200_swiglu_forward_bf16
    with torch.cuda.device(x.device):
        rc = fn(_p(out), _p(x), out.numel(), H, _stream_ptr(x.device))
    _lib.check(rc, "glu_forward")
    return out


def linear_glu_forward(x: torch.Tensor, weight: torch.Tensor, scales: torch.Tensor, policy, kind: int,
                       bias: torch.Tensor | None = None, out: torch.Tensor | None = None,
                       gate_up: torch.Tensor | None = None) -> torch.Tensor:
    """fc_gate_up_->forward followed by geglu_/swiglu_->forward (Gemma.Block.ixx:347-348) as one call: weight is
    the gate|up Linear [2H, K]; `gate_up` is that Linear's own [M, 2H] output tensor, used only when the fused
    kernel does not take the shape."""
    K = x.shape[-1]
    M = x.numel() // K
    H = weight.shape[0] // 2
    if out is None:
        out = torch.empty((*x.shape[:-1], H), dtype=torch.bfloat16, device=x.device)
    if gate_up is None:
        gate_up = torch.empty((M, 2 * H), dtype=torch.bfloat16, device=x.device)
    L = _lib.lib()
    st = _stream_ptr(x.device)
    with torch.cuda.device(x.device):
        if isinstance(policy, PerChannelFp8):
            rc = L.milab200_w8a16_gemm_glu(_p(out), _p(gate_up), _p(x), _p(weight), _p(scales), _p(bias), M, K, H, kind, st)
        else:
            rc = L.milab200_fp4a16_gemm_glu(_p(out), _p(gate_up), _p(x), _p(weight), _p(scales), _p(bias), M, K, H,
                                            policy.kQuantizationGroupSize, kind, st)
    _lib.check(rc, "linear_glu_forward")
    return out


def rmsnorm_linear_glu_forward(x: torch.Tensor, norm_weight: torch.Tensor | None, norm_bias: torch.Tensor | None, eps: float,
                               weight_offset: float, weight: torch.Tensor, scales: torch.Tensor, policy, kind: int,
                               bias: torch.Tensor | None = None, out: torch.Tensor | None = None,
                               gate_up: torch.Tensor | None = None, scratch: torch.Tensor | None = None) -> torch.Tensor:
    """ln_2 -> fc_gate_up -> geglu / swiglu (Gemma.Block.ixx:209-210,347-349; Llama.Block.ixx:883) as one call: `weight` is
    the gate|up Linear [2H, K]; `gate_up` [M, 2H] and `scratch` [M, K] are the intermediate tensors of the reference's
    three-kernel sequence, touched only when the fused decode routes do not take the shape."""
    K = x.shape[-1]
    M = x.numel() // K
    H = weight.shape[0] // 2
    if out is None:
        out = torch.empty((*x.shape[:-1], H), dtype=torch.bfloat16, device=x.device)
    if gate_up is None:
        gate_up = torch.empty((M, 2 * H), dtype=torch.bfloat16, device=x.device)
    if scratch is None:
        scratch = torch.empty((M, K), dtype=torch.bfloat16, device=x.device)
    L = _lib.lib()
    st = _stream_ptr(x.device)
    with torch.cuda.device(x.device):
        if isinstance(policy, PerChannelFp8):
            rc = L.milab200_rmsnorm_w8a16_gemm_glu(_p(out), _p(gate_up), _p(scratch), _p(x), _p(norm_weight), _p(norm_bias), eps,
                                                   weight_offset, _p(weight), _p(scales), _p(bias), M, K, H, kind, st)
        else:
            rc = L.milab200_rmsnorm_fp4a16_gemm_glu(_p(out), _p(gate_up), _p(scratch), _p(x), _p(norm_weight), _p(norm_bias), eps,
                                                    weight_offset, _p(weight), _p(scales), _p(bias), M, K, H,
                                                    policy.kQuantizationGroupSize, kind, st)
    _lib.check(rc, "rmsnorm_linear_glu_forward")
    return out


# ---- the component -------------------------------------------------------------------------

class Linear:
    """Linear<Cuda, BF16, TWeightQuant> (Linear.ixx:82).  Bias-free or biased, inference only."""

    def __init__(self, name: str, config: LinearConfig, device: str | torch.device = "cuda:0",
                 weight_quant=PerGroupFp4()):
        if not isinstance(weight_quant, (PerChannelFp8, PerGroupFp4)):
            raise InvalidArgument("Linear: only the quantized policies are on this path")
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise InvalidArgument("Linear: device type mismatch")           # Linear.ixx:132
        if not torch.cuda.is_available():
            raise MilaB200Error("Linear: no CUDA device — this path has no CPU fallback")
        _lib.lib()                                                            # fail loudly if the .so is missing
        self._name, self._config, self._policy = name, config, weight_quant
        self._built = False
        self.weight_ = self.weight_scales_ = self.bias_ = self.output_ = None
        self._weight_installed = self._output_installed = False
        self._staging: torch.Tensor | None = None

    # -- introspection
    def getName(self) -> str: return self._name
    def isBuilt(self) -> bool: return self._built
    def hasBias(self) -> bool: return self._config.hasBias()
    def getParameterNames(self):
        """Linear::getParameterNames (Linear.ixx:275-283): {"weight"} or {"weight", "bias"}.  The scales are deliberately
        absent — they travel only through saveFlatTensors / loadParameter("weight_scale") so that the archive's blob
        count stays invariant (Linear.ixx:362-364)."""
        return ["weight"] + (["bias"] if self.hasBias() else [])

    def _weight_shape(self):
        N, K = self._config.getOutputFeatures(), self._config.getInputFeatures()
        return (N, K) if isinstance(self._policy, PerChannelFp8) else (N, K // 2)      # Linear.ixx:1032-1046

    def _scale_shape(self):
        N, K = self._config.getOutputFeatures(), self._config.getInputFeatures()
        if isinstance(self._policy, PerChannelFp8):
            return (N,)
        return (N, K // self._policy.kQuantizationGroupSize)                            # Linear.ixx:1047-1053

    def build(self, input_shape) -> None:
        """onBuilding (Linear.ixx:842-887): allocate weight_/weight_scales_/bias_/output_."""
        input_shape = tuple(int(d) for d in input_shape)
        K, N = self._config.getInputFeatures(), self._config.getOutputFeatures()
        if len(input_shape) < 1 or input_shape[-1] != K:
            raise InvalidArgument(f"Linear '{self._name}': build input shape {input_shape} does not end in {K}")
        if isinstance(self._policy, PerGroupFp4):
            g = self._policy.kQuantizationGroupSize
            if g not in (64, 128):
                raise MilaB200Error(f"unsupported group_size={g}")
            if K % g != 0:
                raise InvalidArgument(f"Linear '{self._name}': in_features {K} not divisible by group size {g}")
        dev = self._device
        if not self._weight_installed:
            self.weight_ = torch.zeros(self._weight_shape(), dtype=torch.uint8, device=dev)
            self.weight_scales_ = torch.zeros(self._scale_shape(), dtype=torch.float32, device=dev)
        if self.hasBias():
            self.bias_ = torch.zeros((N,), dtype=torch.bfloat16, device=dev)
        if not self._output_installed:
            self.output_ = torch.empty((*input_shape[:-1], N), dtype=torch.bfloat16, device=dev)
        self._leading_shape = input_shape
        self._built = True

    def getRequiredMemory(self) -> int:
        """MemoryStats footprint (Linear.ixx:692-833): weight + scales (+ bias) bytes."""
        n = 1
        for d in self._weight_shape(): n *= d
        s = 4
        for d in self._scale_shape(): s *= d
        return n + s + (2 * self._config.getOutputFeatures() if self.hasBias() else 0)

    # -- parameters
    def loadParameter(self, name: str, blob: TensorBlob) -> None:
        """Linear.ixx:529-600.  BF16 blob -> quantize on load; storage-dtype blob -> raw copy."""
        if not self._built:
            raise MilaB200Error("Linear must be built before loading parameters")
        N, K = self._config.getOutputFeatures(), self._config.getInputFeatures()
        if name == "weight":
            storage = self._policy.kStorageDtype
            if blob.dtype == storage:
                if tuple(blob.shape) != tuple(self.weight_.shape):
                    raise InvalidArgument(f"Linear '{self._name}': packed weight shape {tuple(blob.shape)} != "
                                          f"{tuple(self.weight_.shape)}")
                self.weight_.copy_(blob.data.view(torch.uint8).reshape(self.weight_.shape), non_blocking=True)
                self._note_weights_written()
                return
            if blob.dtype != "BF16":
                raise InvalidArgument(f"Linear '{self._name}': weight blob dtype {blob.dtype} is neither "
                                      f"{storage} nor BF16")
            if tuple(blob.shape) != (N, K):                                   # CudaLinearOp.Quantize.ixx:67-73
                raise InvalidArgument(f"quantize - shape mismatch: expected [{N},{K}], got {list(blob.shape)}")
            src = blob.data.view(torch.bfloat16).reshape(N, K)
            if self._staging is None or self._staging.numel() < N * K:        # context scratch, grow-only
                self._staging = torch.empty((N * K,), dtype=torch.bfloat16, device=self._device)
            L = _lib.lib()
            with torch.cuda.device(self._device):
                if isinstance(self._policy, PerChannelFp8):
                    rc = L.milab200_quantize_fp8_per_channel(_p(src), _p(self.weight_), _p(self.weight_scales_),
                                                             N, K, _p(self._staging), _stream_ptr(self._device))
                else:
                    rc = L.milab200_quantize_fp4_per_group(_p(src), _p(self.weight_), _p(self.weight_scales_),
                                                           N, K, self._policy.kQuantizationGroupSize,
                                                           _p(self._staging), _stream_ptr(self._device))
            _lib.check(rc, f"Linear '{self._name}' quantize")
            # the host blob may be reused by the caller right after its own synchronize()
            # (Gemma.ixx:515-521); keep a reference until then is the caller's contract.
        elif name == "weight_scale":
            if tuple(blob.shape) != tuple(self.weight_scales_.shape):
                raise InvalidArgument(f"Linear '{self._name}': weight_scale shape mismatch")
            self.weight_scales_.copy_(blob.data.view(torch.float32).reshape(self.weight_scales_.shape),
                                      non_blocking=True)
        elif name == "bias":
            if not self.hasBias():
                return
            if tuple(blob.shape) != (N,):
                raise InvalidArgument(f"Linear '{self._name}': bias shape mismatch")
            self.bias_.copy_(blob.data.view(torch.bfloat16).reshape(N), non_blocking=True)
        else:
            raise InvalidArgument(f"Linear '{self._name}': unknown parameter '{name}' "
                                  "(expected 'weight', 'weight_scale' or 'bias')")

    def installSharedWeight(self, shared_weight: torch.Tensor, shared_scales: torch.Tensor | None = None) -> None:
        """Linear.ixx:614-680.  Quantized instantiations need (weight, scales)."""
        if shared_scales is None:
            raise LogicError(f"Linear '{self._name}': a quantized weight requires its scales "
                             "(use installSharedWeight(weight, scales))")
        if tuple(shared_weight.shape) != self._weight_shape() or tuple(shared_scales.shape) != self._scale_shape():
            raise InvalidArgument(f"Linear '{self._name}': shared weight/scales shape mismatch")
        self.weight_, self.weight_scales_ = shared_weight.view(torch.uint8), shared_scales
        self._weight_installed = True
        self._note_weights_written()          # the tied table may have been produced by kernels still in flight

    def _note_weights_written(self) -> None:
        """Weight bytes were (or may have been) written by kernels that are not this library's quantizers: the next
        decode launch on this device takes plain stream order instead of prefetching weights early (mila_b200_linear.h,
        'Stream order')."""
        with torch.cuda.device(self._device):
            _lib.lib().milab200_note_weights_written()

    def installSharedOutput(self, output: torch.Tensor) -> None:
        """Linear.ixx:682-690: write into (a prefix of) a larger shared slot."""
        if output.dtype != torch.bfloat16 or not output.is_contiguous():
            raise LogicError(f"Linear '{self._name}': shared output must be contiguous BF16")
        self.output_, self._output_installed = output, True

    def saveFlatTensors(self, prefix: str) -> dict:
        """Linear.ixx:370-400: `<prefix>.weight` (+ `.weight_scale`, `.bias`) as host tensors."""
        out = {}
        if not self._weight_installed:
            out[prefix + ".weight"] = self.weight_.cpu()
            out[prefix + ".weight_scale"] = self.weight_scales_.cpu()
        if self.bias_ is not None:
            out[prefix + ".bias"] = self.bias_.cpu()
        return out

    # -- compute
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Linear.ixx:159-190 + CudaLinearOp::forward."""
        if not self._built:
            raise MilaB200Error("Linear must be built before calling forward.")
        K, N = self._config.getInputFeatures(), self._config.getOutputFeatures()
        if x.dim() < 1:
            raise InvalidArgument("Linear: input must have rank >= 1")
        if x.shape[-1] != K:
            raise InvalidArgument(f"Linear: input feature dimension {x.shape[-1]} != {K}")
        if x.dtype != torch.bfloat16 or not x.is_cuda:
            raise InvalidArgument("Linear: input must be a CUDA BF16 tensor")
        x = x.contiguous()
        M = x.numel() // K
        need = M * N
        if self.output_ is None or self.output_.numel() < need:
            if self._output_installed:
                raise LogicError(f"Linear '{self._name}': shared output slot too small")
            self.output_ = torch.empty((need,), dtype=torch.bfloat16, device=self._device)
        out_view = self.output_.view(-1)[:need].view(*x.shape[:-1], N)
        linear_forward(x, self.weight_, self.weight_scales_, self._policy, self.bias_, out_view)
        return out_view

    def backward(self, *_a, **_k):
        raise LogicError("Linear::backward: not supported on quantized weight paths")   # Linear.ixx:225-228

    def synchronize(self) -> None:
        torch.cuda.current_stream(self._device).synchronize()
