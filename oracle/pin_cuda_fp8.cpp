// pin_cuda_fp8.cpp — exposes the CUDA TOOLKIT's own host implementation of the E4M3
// conversion (cuda_fp8.hpp, __nv_cvt_float_to_fp8 — the function Mila's quantizer calls via
// `__nv_fp8_e4m3( float )`, LIN/Kernels/Quantization/CudaFp8WeightQuantization.cu:120) so the
// oracle's own restatement (oracle_f32_to_e4m3) can be pinned against it.
// Test infrastructure only.  The toolkit header is third-party (NVIDIA), not reference code.
// Build: g++ -O2 -fPIC -shared -I/usr/local/cuda/include pin_cuda_fp8.cpp -o libpin_cuda_fp8.so
#include <cstdint>
#include <cstring>
#include <cuda_fp8.h>

extern "C" {

uint8_t pin_nv_f32_to_e4m3_satfinite(float x)
{
    return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
}

float pin_nv_e4m3_to_f32(uint8_t b)
{
    __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3);
    // half -> float by hand (host, no cuda_fp16 runtime dependency)
    uint32_t sign = (uint32_t)(h.x & 0x8000u) << 16;
    uint32_t e = (h.x >> 10) & 0x1Fu, m = h.x & 0x3FFu, out;
    if (e == 0x1F) out = sign | 0x7F800000u | (m << 13);
    else if (e == 0) {
        if (m == 0) out = sign;
        else { int s = 0; while (!(m & 0x400u)) { m <<= 1; ++s; } m &= 0x3FFu;
               out = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | (m << 13); }
    } else out = sign | ((e - 15 + 127) << 23) | (m << 13);
    float f; std::memcpy(&f, &out, 4); return f;
}

// Compare the oracle restatement against the toolkit over a strided sweep of all float bit
// patterns: bits = start, start+stride, ...  Returns the number of mismatches and stores the
// first mismatching bit pattern.
uint8_t oracle_f32_to_e4m3(float);   // from mila_oracle.c (linked in)
uint64_t pin_sweep_e4m3(uint64_t start, uint64_t stride, uint64_t count, uint32_t* first_bad)
{
    uint64_t bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        uint32_t bits = (uint32_t)(start + i * stride);
        float f; std::memcpy(&f, &bits, 4);
        uint8_t a = pin_nv_f32_to_e4m3_satfinite(f);
        uint8_t b = oracle_f32_to_e4m3(f);
        if (a != b) { if (!bad && first_bad) *first_bad = bits; ++bad; }
    }
    return bad;
}

}
