"""ctypes/numpy front-end of the C oracle (oracle/mila_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  All arrays are numpy; BF16 is carried as
uint16 bit patterns, FP8/FP4 storage as uint8.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None
_PIN = None
_REF = None


def build(force: bool = False) -> None:
    """Compile the C oracle (and the toolkit pin helper) with gcc/g++ via oracle/Makefile.
    When /root/reference is present (the build container) also compile the reference's own
    kernels into oracle/_ref/ — building the checker is not using it."""
    need = force or not (_HERE / "libmila_oracle.so").exists() or not (_HERE / "libpin_cuda_fp8.so").exists() \
        or (_HERE / "libmila_oracle.so").stat().st_mtime < (_HERE / "mila_oracle.c").stat().st_mtime
    if need:
        subprocess.run(["make", "-C", str(_HERE), "all"], check=True, capture_output=True)
    ref_root = Path(os.environ.get("MILA_REF_ROOT", "/root/reference"))
    ref_so = _HERE / "_ref" / "libmila_ref_linear.so"
    if ref_root.exists() and (force or not ref_so.exists()
                              or ref_so.stat().st_mtime < (_HERE / "ref_shim.cu").stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "ref", f"REF_ROOT={ref_root}"], check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(str(_HERE / "libmila_oracle.so"))
        c = ctypes
        P = c.c_void_p
        L.oracle_bf16_to_f32.restype = c.c_float; L.oracle_bf16_to_f32.argtypes = [c.c_uint16]
        L.oracle_f32_to_bf16.restype = c.c_uint16; L.oracle_f32_to_bf16.argtypes = [c.c_float]
        L.oracle_f32_to_e4m3.restype = c.c_uint8; L.oracle_f32_to_e4m3.argtypes = [c.c_float]
        L.oracle_e4m3_to_f32.restype = c.c_float; L.oracle_e4m3_to_f32.argtypes = [c.c_uint8]
        L.oracle_f32_to_e2m1.restype = c.c_uint8; L.oracle_f32_to_e2m1.argtypes = [c.c_float]
        L.oracle_e2m1_to_f32.restype = c.c_float; L.oracle_e2m1_to_f32.argtypes = [c.c_uint8]
        L.oracle_quantize_fp8_per_channel.restype = None
        L.oracle_quantize_fp8_per_channel.argtypes = [P, P, P, c.c_int64, c.c_int64]
        L.oracle_quantize_fp4_per_group.restype = c.c_int
        L.oracle_quantize_fp4_per_group.argtypes = [P, P, P, c.c_int64, c.c_int64, c.c_int]
        L.oracle_dequant_fp8.restype = None
        L.oracle_dequant_fp8.argtypes = [P, P, P, c.c_int64, c.c_int64]
        L.oracle_dequant_fp4.restype = c.c_int
        L.oracle_dequant_fp4.argtypes = [P, P, P, c.c_int64, c.c_int64, c.c_int]
        L.oracle_rmsnorm_forward_bf16.restype = None
        L.oracle_rmsnorm_forward_bf16.argtypes = [P, P, P, P, P, c.c_int64, c.c_int64, c.c_float, c.c_float]
        L.oracle_w4a16_int4_forward.restype = None
        L.oracle_w4a16_int4_forward.argtypes = [P, P, P, P, P, P, P, c.c_int64, c.c_int64, c.c_int64, c.c_int]
        L.oracle_linear_forward_bf16.restype = None
        L.oracle_linear_forward_bf16.argtypes = [P, P, P, P, P, c.c_int64, c.c_int64, c.c_int64]
        for name in ("oracle_cpu_linear_forward_naive", "oracle_cpu_linear_forward_unrolled",
                     "oracle_cpu_linear_forward", "oracle_ctx_linear_forward_f32acc"):
            f = getattr(L, name); f.restype = None
            f.argtypes = [P, P, P, P, c.c_int64, c.c_int64, c.c_int64]
        L.oracle_compute_fp8_weight_scale.restype = c.c_float
        L.oracle_compute_fp8_weight_scale.argtypes = [P, c.c_int64]
        L.oracle_fp4_dequantize_to_fp8.restype = c.c_int
        L.oracle_fp4_dequantize_to_fp8.argtypes = [P, P, c.c_float, P, c.c_int64, c.c_int64, c.c_int]
        L.oracle_fp4_dequantize_to_bf16.restype = c.c_int
        L.oracle_fp4_dequantize_to_bf16.argtypes = [P, P, P, c.c_int64, c.c_int64, c.c_int]
        L.oracle_fp8_dequantize_to_bf16.restype = None
        L.oracle_fp8_dequantize_to_bf16.argtypes = [P, P, P, c.c_int64, c.c_int64]
        L.oracle_quantize_bf16_to_fp8_per_token.restype = None
        L.oracle_quantize_bf16_to_fp8_per_token.argtypes = [P, P, P, c.c_int64, c.c_int64]
        L.oracle_token_embedding_qfp8.restype = None
        L.oracle_token_embedding_qfp8.argtypes = [P, P, P, P, c.c_int64, c.c_int64]
        L.oracle_glu_forward_bf16.restype = c.c_int
        L.oracle_glu_forward_bf16.argtypes = [P, P, c.c_int64, c.c_int64, c.c_int]
        _LIB = L
    return _LIB


def pin_lib() -> ctypes.CDLL:
    """The CUDA toolkit's own host E4M3 conversion (cuda_fp8.hpp), for pinning."""
    global _PIN
    if _PIN is None:
        build()
        L = ctypes.CDLL(str(_HERE / "libpin_cuda_fp8.so"))
        L.pin_nv_f32_to_e4m3_satfinite.restype = ctypes.c_uint8
        L.pin_nv_f32_to_e4m3_satfinite.argtypes = [ctypes.c_float]
        L.pin_nv_e4m3_to_f32.restype = ctypes.c_float
        L.pin_nv_e4m3_to_f32.argtypes = [ctypes.c_uint8]
        L.pin_sweep_e4m3.restype = ctypes.c_uint64
        L.pin_sweep_e4m3.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                     ctypes.POINTER(ctypes.c_uint32)]
        _PIN = L
    return _PIN


def ref_lib_path() -> Path:
    return _HERE / "_ref" / "libmila_ref_linear.so"


def ref_lib() -> ctypes.CDLL:
    """The reference's own CUDA kernels (compiled unmodified) — needs a GPU to call."""
    global _REF
    if _REF is None:
        p = ref_lib_path()
        if not p.exists():
            raise FileNotFoundError(f"{p} missing: run `make -C oracle ref` where /root/reference exists")
        _REF = ctypes.CDLL(str(p))
    return _REF


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


# ---- dtype helpers ---------------------------------------------------------------------

def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """RNE FP32 -> BF16 bit patterns (vectorised; same rule as oracle_f32_to_bf16)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + (0x7FFF + ((u >> 16) & 1))) >> 16).astype(np.uint16)
    r[nan] = 0x7FFF
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


# ---- quantizers ------------------------------------------------------------------------

def quantize_fp8_per_channel(w_bf16: np.ndarray):
    w = _c(w_bf16, np.uint16); N, K = w.shape
    q = np.empty((N, K), np.uint8); s = np.empty((N,), np.float32)
    lib().oracle_quantize_fp8_per_channel(_p(w), _p(q), _p(s), N, K)
    return q, s


def quantize_fp4_per_group(w_bf16: np.ndarray, group_size: int = 128):
    w = _c(w_bf16, np.uint16); N, K = w.shape
    if group_size not in (64, 128):
        raise RuntimeError(f"unsupported group_size={group_size}")
    if K % group_size:
        raise ValueError("in_features must be divisible by group_size")
    q = np.empty((N, K // 2), np.uint8); s = np.empty((N, K // group_size), np.float32)
    rc = lib().oracle_quantize_fp4_per_group(_p(w), _p(q), _p(s), N, K, group_size)
    assert rc == 0
    return q, s


def dequant_fp8(q: np.ndarray, s: np.ndarray) -> np.ndarray:
    q = _c(q, np.uint8); s = _c(s, np.float32); N, K = q.shape
    out = np.empty((N, K), np.float32)
    lib().oracle_dequant_fp8(_p(q), _p(s), _p(out), N, K)
    return out


def dequant_fp4(q: np.ndarray, s: np.ndarray, group_size: int = 128) -> np.ndarray:
    q = _c(q, np.uint8); s = _c(s, np.float32); N, K2 = q.shape; K = K2 * 2
    out = np.empty((N, K), np.float32)
    rc = lib().oracle_dequant_fp4(_p(q), _p(s), _p(out), N, K, group_size)
    assert rc == 0
    return out


# ---- forward ---------------------------------------------------------------------------

def linear_forward_bf16(x_bf16: np.ndarray, wf: np.ndarray, bias_bf16: np.ndarray | None = None):
    """dequantise-then-FP32-GEMM oracle.  Returns (y_bf16_bits [M,N], y_f32 [M,N])."""
    x = _c(x_bf16, np.uint16); wf = _c(wf, np.float32)
    M, K = x.shape; N = wf.shape[0]; assert wf.shape[1] == K
    b = None if bias_bf16 is None else _c(bias_bf16, np.uint16)
    y = np.empty((M, N), np.uint16); yf = np.empty((M, N), np.float32)
    lib().oracle_linear_forward_bf16(_p(x), _p(wf), _p(b), _p(y), _p(yf), M, K, N)
    return y, yf


def linear_forward_fp8(x_bf16, q, s, bias_bf16=None):
    return linear_forward_bf16(x_bf16, dequant_fp8(q, s), bias_bf16)


def linear_forward_fp4(x_bf16, q, s, group_size=128, bias_bf16=None):
    return linear_forward_bf16(x_bf16, dequant_fp4(q, s, group_size), bias_bf16)


def rmsnorm_forward_bf16(x_bf16, weight_bf16=None, bias_bf16=None, eps=1e-6, weight_offset=0.0):
    """BF16 RMSNorm over the last axis (RmsNorm.Bf16.cu:19-73 restated): returns (BF16 bits [M,K], rstd f32 [M])."""
    x = _c(x_bf16, np.uint16); M, K = x.shape
    w = None if weight_bf16 is None else _c(weight_bf16, np.uint16)
    b = None if bias_bf16 is None else _c(bias_bf16, np.uint16)
    y = np.empty((M, K), np.uint16); r = np.empty((M,), np.float32)
    lib().oracle_rmsnorm_forward_bf16(_p(x), _p(w), _p(b), _p(y), _p(r), M, K, ctypes.c_float(eps), ctypes.c_float(weight_offset))
    return y, r


def w4a16_int4_forward(x_bf16, w_packed, scales, zero_points=None, bias_bf16=None, group_size=128):
    """PerGroupInt4 forward (CudaW4A16Gemm.cu:88-197 restated): returns (BF16 bits [M,N], FP32 [M,N])."""
    x = _c(x_bf16, np.uint16); w = _c(w_packed, np.uint8); s = _c(scales, np.float32)
    M, K = x.shape; N = w.shape[0]
    z = None if zero_points is None else _c(zero_points, np.uint8)
    b = None if bias_bf16 is None else _c(bias_bf16, np.uint16)
    y = np.empty((M, N), np.uint16); yf = np.empty((M, N), np.float32)
    lib().oracle_w4a16_int4_forward(_p(x), _p(w), _p(s), _p(z), _p(b), _p(y), _p(yf), M, K, N, group_size)
    return y, yf


def token_embedding_qfp8(ids: np.ndarray, w8: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """FP8 tied-table gather-dequant -> BF16 bits [len(ids), C] (TokenEmbedding.Fp8.cu:34-66)."""
    ids = _c(ids, np.int32).reshape(-1); w8 = _c(w8, np.uint8); scales = _c(scales, np.float32)
    C = w8.shape[1]
    y = np.empty((ids.size, C), np.uint16)
    lib().oracle_token_embedding_qfp8(_p(ids), _p(w8), _p(scales), _p(y), ids.size, C)
    return y


def glu_forward_bf16(x_bf16: np.ndarray, kind: int) -> np.ndarray:
    """GeGLU (kind 1) / SwiGLU (kind 2) over BF16 bits [tokens, 2H] -> BF16 bits [tokens, H]
    (Geglu.cu:42-61, Swiglu.Bf16.cu:135-230)."""
    x = _c(x_bf16, np.uint16)
    tokens, twoH = x.shape
    y = np.empty((tokens, twoH // 2), np.uint16)
    rc = lib().oracle_glu_forward_bf16(_p(x), _p(y), ctypes.c_int64(tokens), ctypes.c_int64(twoH // 2), int(kind))
    assert rc == 0
    return y


def cpu_linear_forward(X: np.ndarray, W: np.ndarray, B: np.ndarray | None = None, path: str = "auto"):
    """Restated CpuLinearOp::forward (FP32).  path: auto|naive|unrolled (the reference's), context_f32acc (a labelled
    non-reference context row: float accumulator)."""
    X = _c(X, np.float32); W = _c(W, np.float32)
    batch, K = X.shape; N = W.shape[0]
    b = None if B is None else _c(B, np.float32)
    Y = np.zeros((batch, N), np.float32)
    fn = {"auto": lib().oracle_cpu_linear_forward, "naive": lib().oracle_cpu_linear_forward_naive,
          "unrolled": lib().oracle_cpu_linear_forward_unrolled,
          "context_f32acc": lib().oracle_ctx_linear_forward_f32acc}[path]        # labelled context row, not the reference
    fn(_p(X), _p(Y), _p(W), _p(b), batch, K, N)
    return Y


# ---- W4A8 helpers ----------------------------------------------------------------------

def compute_fp8_weight_scale(group_scales: np.ndarray) -> float:
    g = _c(group_scales, np.float32).reshape(-1)
    return float(lib().oracle_compute_fp8_weight_scale(_p(g), g.size))


def fp4_dequantize_to_fp8(q, s, sB: float, group_size=128):
    q = _c(q, np.uint8); s = _c(s, np.float32); N, K2 = q.shape
    out = np.empty((N, K2 * 2), np.uint8)
    rc = lib().oracle_fp4_dequantize_to_fp8(_p(q), _p(s), ctypes.c_float(sB), _p(out), N, K2 * 2, group_size)
    assert rc == 0
    return out


def fp4_dequantize_to_bf16(q, s, group_size=128):
    q = _c(q, np.uint8); s = _c(s, np.float32); N, K2 = q.shape
    out = np.empty((N, K2 * 2), np.uint16)
    rc = lib().oracle_fp4_dequantize_to_bf16(_p(q), _p(s), _p(out), N, K2 * 2, group_size)
    assert rc == 0
    return out


def fp8_dequantize_to_bf16(q, s):
    q = _c(q, np.uint8); s = _c(s, np.float32); N, K = q.shape
    out = np.empty((N, K), np.uint16)
    lib().oracle_fp8_dequantize_to_bf16(_p(q), _p(s), _p(out), N, K)
    return out


def quantize_bf16_to_fp8_per_token(x_bf16):
    x = _c(x_bf16, np.uint16); M, K = x.shape
    x8 = np.empty((M, K), np.uint8); sA = np.empty((M,), np.float32)
    lib().oracle_quantize_bf16_to_fp8_per_token(_p(x), _p(x8), _p(sA), M, K)
    return x8, sA


# ---- the reference's own deterministic fixtures ------------------------------------------

def ref_weight_value(out_index, in_index):
    """weightValue — Tests/Dnn/Components/Linear/Linear.Cuda.cpp:70-75 (vectorised)."""
    h = (np.asarray(out_index, np.int64) * 13 + np.asarray(in_index, np.int64) * 7) % 17
    return (np.float32(0.1) * (h.astype(np.float32) - np.float32(8.0)) / np.float32(17.0)).astype(np.float32)


def ref_bias_value(out_index):
    """biasValue — Linear.Cuda.cpp:77-80."""
    o = np.asarray(out_index, np.int64)
    return (np.float32(0.1) * ((o % 5).astype(np.float32) - np.float32(2.0)) / np.float32(5.0)).astype(np.float32)


def ref_weight_blob(N: int, K: int) -> np.ndarray:
    """BF16 weight blob exactly as the reference tests build it (Linear.Cuda.cpp:666-677)."""
    o, i = np.meshgrid(np.arange(N), np.arange(K), indexing="ij")
    return f32_to_bf16_bits(ref_weight_value(o, i))


def ref_magnitude_rows(M: int, K: int) -> np.ndarray:
    """15-decade row-magnitude activation fixture — Linear.Cuda.cpp:822-833 (FP32 values)."""
    m, k = np.meshgrid(np.arange(M), np.arange(K), indexing="ij")
    row_scale = np.power(np.float32(10.0), (m.astype(np.float32) - np.float32(8.0))).astype(np.float32)
    spread = (((m * 31 + k * 17) % 257).astype(np.float32) / np.float32(128.0) - np.float32(1.0)).astype(np.float32)
    return (row_scale * spread).astype(np.float32)
