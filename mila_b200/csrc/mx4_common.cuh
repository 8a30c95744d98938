// mx4_common.cuh — device helpers of the packed-nibble kind::mxf4 decode path shared by decode_mx4.cu (one launch per
// Linear) and decode_chain.cu (chained launch): the block-scaled MMA wrapper, the TMEM scale-factor store, and the exact
// split of an activation into eight 2-bit digit planes whose nibbles are E2M1 codes (decode_mx4.cu header comment).
#pragma once
#include "common.cuh"

namespace milab200 {
namespace mx4 {

__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t tsfa, uint32_t tsfb, uint32_t accumulate)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
                 "  tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p; }"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb) : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, uint32_t v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(taddr), "r"(v) : "memory");
}

// One BF16 activation scaled into [-2^16, 2^16) -> a word whose nibble p is the E2M1 code of 2-bit digit p.
__device__ __forceinline__ uint32_t digit_codes(float x, float scale)
{
    const int w = __float2int_rn(x * scale);
    const uint32_t u = (uint32_t)abs(w);
    uint32_t t = (u | (u << 8)) & 0x00FF00FFu;
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    t = (t | (t << 2)) & 0x33333333u;                       // nibble p = digit p (0..3)
    const uint32_t c3 = t & (t >> 1) & 0x11111111u;         // digit == 3
    uint32_t code = (t << 1) - c3;                          // 0,1,2,3 -> E2M1 codes 0,2,4,5 (0, 1, 2, 3)
    if (w < 0) code |= 0x88888888u;
    return code;
}

// 8 elements x 8 digit nibbles -> 8 plane words (nibble i of word p = digit p of element i).
__device__ __forceinline__ void transpose_nibbles_8x8(uint32_t (&W)[8])
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t a = W[i], b = W[i + 4];
        W[i] = __byte_perm(a, b, 0x5410); W[i + 4] = __byte_perm(a, b, 0x7632);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = (q & 1) + ((q >> 1) << 2);            // 0, 1, 4, 5
        const uint32_t a = W[i], b = W[i + 2];
        W[i] = __byte_perm(a, b, 0x6240); W[i + 2] = __byte_perm(a, b, 0x7351);
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const uint32_t a = W[i], b = W[i + 1];
        W[i] = (a & 0x0F0F0F0Fu) | ((b & 0x0F0F0F0Fu) << 4);
        W[i + 1] = ((a >> 4) & 0x0F0F0F0Fu) | (b & 0xF0F0F0F0u);
    }
}

}  // namespace mx4
}  // namespace milab200
