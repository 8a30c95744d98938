"""CPU model of the decode kernel's data movement (mila_b200/csrc/gemv.cu).

The MMA kernel permutes K (each lane holds 16 contiguous weight bytes of its row; the tensor
core's logical k index is mapped to those physical positions) and stages activations in
B-fragment order.  This test re-states the index formulas in numpy and checks that the
contraction they describe equals a plain matmul — it guards the layout algebra on a machine
with no GPU.  (The GPU parity tests check the real kernel.)
"""
import numpy as np
import pytest

FMTS = {
    # name: (KT, STEP, Q, elements per 32-bit word)
    "fp8": (16, 64, 2, 4),
    "fp4g128": (32, 128, 4, 8),
    "fp4g64": (16, 64, 2, 8),
}


def mma_m16n8k16(A_frag, B_frag):
    """A_frag[lane] = [a0..a7] (PTX m16n8k16 .row layout), B_frag[lane] = [b0..b3] (.col).
    Returns D_frag[lane] = [d0..d3]."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        a = A_frag[lane]
        A[g, 2 * t], A[g, 2 * t + 1] = a[0], a[1]
        A[g + 8, 2 * t], A[g + 8, 2 * t + 1] = a[2], a[3]
        A[g, 2 * t + 8], A[g, 2 * t + 9] = a[4], a[5]
        A[g + 8, 2 * t + 8], A[g + 8, 2 * t + 9] = a[6], a[7]
        b = B_frag[lane]
        B[2 * t, g], B[2 * t + 1, g] = b[0], b[1]
        B[2 * t + 8, g], B[2 * t + 9, g] = b[2], b[3]
    D = A @ B
    out = np.zeros((32, 4))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane] = [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]
    return out


@pytest.mark.parametrize("fmt", list(FMTS))
@pytest.mark.parametrize("M", [1, 3, 8])
def test_step_contraction_matches_matmul(fmt, M):
    KT, STEP, Q, epw = FMTS[fmt]
    rng = np.random.default_rng(0)
    W = rng.integers(-3, 4, size=(16, STEP)).astype(np.float64)       # one row tile, one step
    X = rng.integers(-3, 4, size=(M, STEP)).astype(np.float64)

    # staging: element k of token m -> xs[(q*M + m)*4 + tt] (8 halves)
    xs = np.zeros((Q * M * 4, 8))
    for m in range(M):
        for c in range(STEP // 8):
            k = c * 8
            tt, q = k // KT, (k % KT) // 8
            xs[(q * M + m) * 4 + tt] = X[m, k:k + 8]

    acc = np.zeros((32, 4))
    words_per_thread = KT // epw
    for wq in range(words_per_thread):             # 32-bit word index inside the lane's 16 bytes
        # pairs of the word: element offsets (2b, 2b+1), b = 0..epw/2-1
        for c in range(epw // 4):                  # MMAs per word: fp8 -> 1, fp4 -> 2
            A_frag = np.zeros((32, 8)); B_frag = np.zeros((32, 4))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                kbase = t * KT + wq * epw + 4 * c   # 4 consecutive k per MMA per lane
                lo, hi = W[g], W[g + 8]
                A_frag[lane] = [lo[kbase], lo[kbase + 1], hi[kbase], hi[kbase + 1],
                                lo[kbase + 2], lo[kbase + 3], hi[kbase + 2], hi[kbase + 3]]
                if g < M:
                    # activation chunk q holds k offsets t*KT + 8q .. +8
                    off = wq * epw + 4 * c
                    q, h = off // 8, off % 8
                    row = xs[(q * M + g) * 4 + t]
                    B_frag[lane] = row[h:h + 4]
            acc += mma_m16n8k16(A_frag, B_frag)

    ref = W @ X.T                                   # [16 rows, M tokens]
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for j in range(4):
            row = g + 8 * (j >> 1); tok = 2 * t + (j & 1)
            if tok < M:
                assert acc[lane, j] == ref[row, tok]
            else:
                assert acc[lane, j] == 0
