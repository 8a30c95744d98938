"""Shared synthetic-input generators and error metrics for the parity tests."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O


def xavier_weights_bf16(N: int, K: int, seed: int = 1234) -> np.ndarray:
    """BF16 bits of randn * 1/sqrt(K) (what Linear uses for init, Linear.ixx:1060)."""
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((N, K), dtype=np.float32) / np.float32(np.sqrt(K))).astype(np.float32)
    return O.f32_to_bf16_bits(w)


def activations_bf16(M: int, K: int, seed: int = 99) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return O.f32_to_bf16_bits(rng.standard_normal((M, K), dtype=np.float32))


def adversarial_weights_bf16(N: int, K: int, group: int = 128, seed: int = 4321) -> np.ndarray:
    """Corner-case pack for the quantizers: zero rows/groups, threshold ties, +-0, outliers,
    BF16 subnormals, NaN, Inf.  N >= 12 rows, K a multiple of `group`."""
    assert N >= 12 and K % group == 0
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((N, K), dtype=np.float32) * np.float32(0.05)).astype(np.float32)
    w[0, :] = 0.0                                            # all-zero row  -> scale 1
    w[1, :group] = 0.0                                       # all-zero group
    # a group sitting exactly on the 7 E2M1 thresholds (x absmax 6 -> scale 1)
    thr = np.array([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0], np.float32)
    g = np.zeros(group, np.float32); g[0] = 6.0
    g[1:8] = thr; g[8:15] = -thr
    g[15] = -0.0; g[16] = 0.0; g[17] = -1e-30; g[18] = 1e-30
    w[2, :group] = g
    # same ties with a non-trivial scale (absmax 3 -> scale 0.5)
    w[3, :group] = g * np.float32(0.5)
    w[4, 5] = 1.0e4                                          # one huge outlier
    w[5, :] = w[5, :] * np.float32(1e-38)                    # subnormal / tiny magnitudes
    w[6, 3] = np.nan                                         # NaN inside a row
    w[7, :group] = np.nan                                    # an all-NaN group
    w[8, 7] = np.inf                                         # +Inf
    w[9, 9] = -np.inf
    w[10, :] = np.float32(448.0) * np.sign(w[10, :])         # saturating magnitudes
    w[11, :] = np.float32(-0.0)
    bits = O.f32_to_bf16_bits(w)
    # keep NaN payloads canonical but exercise a negative NaN too
    bits[6, 4] = 0xFFC0
    return bits


def rel_err_rowabs(y: np.ndarray, ref: np.ndarray) -> float:
    """SURVEY §8d gate: max_i |y_i - ref_i| / max(|ref_i|, 1e-2 * row_absmax)."""
    y = np.asarray(y, np.float64); ref = np.asarray(ref, np.float64)
    row_abs = np.max(np.abs(ref), axis=-1, keepdims=True)
    den = np.maximum(np.abs(ref), 1e-2 * row_abs)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(y - ref) / den))
