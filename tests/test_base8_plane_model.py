"""Host model of the activation format of the 8-token kind::mxf4 decode variant (decode_mx4.cu: base8_planes_x8):
signed base-8 digit planes.  Restates the device bit operations one for one (including PRMT as an 8-entry byte table)
and checks them against the definition:  u = sum_p 8^p d_p,  d_p in {-3 .. 4},  every d_p an E2M1 number whose 4-bit code
sits in nibble i of plane word p for element i.  The hardware then recombines the planes because plane p's B-side
UE8M0 scale factor is 2^(3p) — the GPU tests check that end to end; this checks the packing on the CPU."""
import numpy as np

BIAS = 3 * ((1 << 18) - 1) // 7            # 0o333333: turns signed digits -3..4 into octal digits 0..7
E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def e2m1_value(code: int) -> float:
    return (-1.0 if code & 8 else 1.0) * E2M1[code & 7]


def prmt(a: int, b: int, sel: int) -> int:
    """PRMT (default mode): output byte i = byte sel.nibble[i] of {a (0-3), b (4-7)}."""
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for i in range(4):
        n = (sel >> (4 * i)) & 0xF
        assert n < 8, "sign-replicate selectors are never produced"
        out |= src[n] << (8 * i)
    return out


def base8_planes_x8(u8):
    """u8: the 8 integers rn(x * 2^(15-E)) of one lane's chunk -> six plane words (decode_mx4.cu, same operations)."""
    A, B = [0] * 4, [0] * 4
    for j in range(4):
        ra, rb = [0, 0], [0, 0]
        for h in range(2):
            o = (u8[2 * j + h] + BIAS) & 0xFFFFFFFF
            t = (o & 0x1FF) | ((o << 7) & 0x01FF0000)
            t = (t & 0x00070007) | ((t & 0x00380038) << 1) | ((t & 0x01C001C0) << 2)
            ra[h] = prmt(0x000A0C0D, 0x06050402, t & 0xFFFF)
            rb[h] = prmt(0x000A0C0D, 0x06050402, (t >> 16) & 0xFFFF)
        A[j] = (ra[0] | (ra[1] << 4)) & 0xFFFFFFFF
        B[j] = (rb[0] | (rb[1] << 4)) & 0xFFFFFFFF
    W = [0] * 6
    for p in range(3):
        sel = p | ((4 + p) << 4)
        W[p] = prmt(prmt(A[0], A[1], sel), prmt(A[2], A[3], sel), 0x5410)
        W[3 + p] = prmt(prmt(B[0], B[1], sel), prmt(B[2], B[3], sel), 0x5410)
    return W


def _check(u8):
    W = base8_planes_x8(u8)
    for i in range(8):
        digits = [e2m1_value((W[p] >> (4 * i)) & 0xF) for p in range(6)]
        assert all(d in (-3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 4.0) for d in digits)
        assert sum((8 ** p) * d for p, d in enumerate(digits)) == u8[i], (u8, i, digits)


def test_digit_range_covers_the_17_bit_fixed_point_magnitude():
    lo, hi = -BIAS, 8 ** 6 - 1 - BIAS                   # representable range of six digits in {-3 .. 4}
    assert lo <= -65535 and hi >= 65535                 # |rn(x 2^(15-E))| <= 65280 for BF16 x with |x| < 2^(E+1)
    assert BIAS == 0o333333


def test_corner_values():
    _check([0, 65535, -65535, 1, -1, 3, -4, 32768])
    _check([4, -3, 8, -8, 36, -28, 0o177777, -0o177777])
    _check([65280, -65280, 255, -255, 256, -256, 4095, -4096])   # BF16 significands at the block maximum


def test_random_values_reconstruct_exactly():
    rng = np.random.default_rng(0)
    for _ in range(3000):
        _check([int(v) for v in rng.integers(-65535, 65536, 8)])


def test_zero_activation_is_the_all_zero_nibble():
    """Padding rows and dead tokens are zero bytes; a zero activation must map to the same bytes (+0, never -0)."""
    assert base8_planes_x8([0] * 8) == [0] * 6
