"""CPU test: the drop-in launcher header (include/mila_b200/Kernels/MilaB200Linear.cuh) really compiles as C++ and links
against libmila_b200_linear.so.  A translation unit calls every launcher with the reference's exact argument types
(K/Linear.cuh:49-81, K/Quantization/*.cuh, K/W8A16Gemm, K/W4A16Gemm, K/Fp8Prefill, Embeddings, Activations, RmsNorm
headers) inside a never-taken branch, so every inline wrapper is instantiated and every C-ABI symbol it forwards to
must resolve at link time; the program then runs (no GPU needed: it only queries the ABI version)."""
import shutil
import subprocess
from pathlib import Path

import pytest

from mila_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent

TU = r'''
#include <cstdio>
#include "mila_b200/Kernels/MilaB200Linear.cuh"
namespace L = Mila::Dnn::Compute::Cuda::Linear;
int main(int argc, char**)
{
    if (argc > 1000) {   // never taken: instantiates every wrapper with Mila's own argument types
        __nv_bfloat16 *y = nullptr; const __nv_bfloat16 *x = nullptr, *b = nullptr; const __nv_fp8_e4m3* w8 = nullptr;
        const uint8_t *w4 = nullptr, *zp = nullptr; const float* sc = nullptr; float* fs = nullptr; float* f32 = nullptr;
        void* stage = nullptr; const void* src = nullptr; const int* ids = nullptr; cudaStream_t st = nullptr; void* ctx = nullptr;
        L::cuda_quantize_fp8_per_channel(src, stage, fs, 1, 1, stage, st);
        L::cuda_quantize_fp4_per_group(src, stage, fs, 1, 128, 128, stage, st);
        L::cuda_matvec_decode_bf16_qfp8(y, x, w8, sc, b, 1, 1, st);
        L::cuda_matvec_decode_bf16_qfp4(y, x, w4, sc, b, 128, 1, 128, st);
        L::cuda_w8a16_gemm(y, x, w8, sc, b, 1, 1, 1, st);
        L::cuda_fp4a16_gemm(y, x, w4, sc, b, 1, 128, 1, 128, st);
        L::cuda_fp4a16_gemm_wmma(y, x, w4, sc, b, 1, 128, 1, 128, st);
        L::cuda_w4a16_gemm(y, x, w4, sc, zp, b, 1, 128, 1, 128, st);
        L::cuda_fp8_dequantize_to_bf16(y, w8, sc, 1, 1, st);
        L::cuda_fp4_dequantize_to_bf16(y, w4, sc, 1, 128, 128, st);
        L::cuda_compute_fp8_weight_scale(fs, sc, 1, st);
        L::cuda_fp4_dequantize_to_fp8((__nv_fp8_e4m3*)stage, w4, sc, sc, 1, 128, 128, st);
        L::cuda_quantize_bf16_to_fp8_per_token((__nv_fp8_e4m3*)stage, fs, x, 1, 1, st);
        L::cuda_fp8_apply_per_token_scales(y, sc, b, 1, 1, st);
        L::cuda_add_bias(y, b, 1, 1, st);
        L::cuda_add_bias(f32, sc, 1, 1, st);
        L::cuda_w8a16_gemm_rowparallel_nccl(y, x, w8, sc, b, 32, 128, 128, ctx, st);
        L::cuda_fp4a16_gemm_rowparallel_nccl(y, x, w4, sc, b, 32, 128, 128, 128, ctx, st);
        L::cuda_rmsnorm_w8a16_gemm(y, y, x, x, b, 1e-6f, 1.0f, w8, sc, b, 1, 128, 128, st);
        L::cuda_rmsnorm_fp4a16_gemm(y, y, x, x, b, 1e-6f, 0.0f, w4, sc, b, 1, 128, 128, 128, st);
        Mila::Dnn::Compute::Cuda::TokenEmbedding::cuda_token_embedding_forward_bf16_qfp8(y, ids, src, sc, 1, 1, 8, st);
        Mila::Dnn::Compute::Cuda::TokenEmbedding::cuda_token_embedding_decode_bf16_qfp8(y, ids, src, sc, 1, 8, st);
        Mila::Dnn::Compute::Cuda::Geglu::cuda_geglu_forward_bf16(y, x, 1, 1, st);
        Mila::Dnn::Compute::Cuda::Swiglu::cuda_swiglu_forward_bf16(y, x, 1, 1, st);
        Mila::Dnn::Compute::Cuda::RmsNorm::cuda_rmsnorm_forward_bf16(y, y, x, x, b, 1, 1, 8, 1e-6f, 0.0f, st);
    }
    std::printf("abi %d\n", milab200_abi_version());
    return 0;
}
'''


def test_drop_in_header_compiles_links_and_runs(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    _lib.lib()
    src = tmp_path / "tu.cu"; exe = tmp_path / "tu"
    src.write_text(TU)
    r = subprocess.run([nvcc, "-std=c++20", "-O0", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                        "-L", str(ROOT / "mila_b200"), "-lmila_b200_linear", "-Xlinker", f"-rpath={ROOT / 'mila_b200'}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "abi 1", (out.stdout, out.stderr)
