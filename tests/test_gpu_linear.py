"""GPU tests of the host mirror, written after Mila/Tests/Dnn/Components/Linear/Linear.Cuda.cpp's
quantized cases (:618-:1427): relational exact pins, save/reload, error conventions."""
import numpy as np
import pytest
import torch

import gpu_util as G
from mila_b200 import _lib
from mila_b200.linear import Linear, LinearConfig, PerChannelFp8, PerGroupFp4, TensorBlob
from oracle import oracle as O

pytestmark = pytest.mark.gpu
K_IN, K_OUT = 64, 32


def _blob(N, K):
    bits = O.ref_weight_blob(N, K)
    return TensorBlob("BF16", (N, K), G.bf16_tensor(bits, "cpu")), bits


def _spread(shape):
    n = int(np.prod(shape))
    v = (np.arange(n, dtype=np.float32) / np.float32(n) * np.float32(2.0) - np.float32(1.0)).reshape(shape)
    return G.bf16_tensor(O.f32_to_bf16_bits(v), "cuda")            # spreadHost, Linear.Cuda.cpp:214-224


def test_fp8_direct_load_equals_shared_weight_install():
    """:649 InstallSharedWeight_PerChannelFp8_MatchesDirectQuantizedLoad — decode outputs EXPECT_EQ."""
    cfg = LinearConfig(K_IN, K_OUT).withBias(False)
    blob, _ = _blob(K_OUT, K_IN)
    direct = Linear("lm_head_direct", cfg, "cuda:0", PerChannelFp8())
    direct.build((1, K_IN)); direct.loadParameter("weight", blob)
    donor = Linear("embedding_like", cfg, "cuda:0", PerChannelFp8())
    donor.build((1, K_IN)); donor.loadParameter("weight", blob)
    tied = Linear("lm_head_tied", cfg, "cuda:0", PerChannelFp8())
    tied.installSharedWeight(donor.weight_, donor.weight_scales_)
    tied.build((1, K_IN))
    x = _spread((1, K_IN))
    assert torch.equal(direct.forward(x), tied.forward(x))
    assert tied.saveFlatTensors("h") == {}                          # borrowed weight is not emitted


def test_install_shared_weight_without_scales_throws():
    """:618,:631 — quantized install needs (weight, scales): std::logic_error."""
    lin = Linear("l", LinearConfig(K_IN, K_OUT).withBias(False), "cuda:0", PerChannelFp8())
    with pytest.raises(_lib.LogicError):
        lin.installSharedWeight(torch.zeros((K_OUT, K_IN), dtype=torch.uint8, device="cuda"))


def test_fp8_emits_weight_and_scales_that_reconstruct():
    """:952 PerChannelFp8_EmitsWeightAndScalesThatReconstructTheWeight."""
    lin = Linear("l", LinearConfig(K_IN, K_OUT).withBias(False), "cuda:0", PerChannelFp8())
    lin.build((1, K_IN)); blob, bits = _blob(K_OUT, K_IN); lin.loadParameter("weight", blob)
    t = lin.saveFlatTensors("l")
    assert t["l.weight"].shape == (K_OUT, K_IN) and t["l.weight"].dtype == torch.uint8
    assert t["l.weight_scale"].shape == (K_OUT,) and t["l.weight_scale"].dtype == torch.float32
    w = O.bf16_bits_to_f32(bits)
    rec = O.dequant_fp8(t["l.weight"].numpy(), t["l.weight_scale"].numpy())
    assert np.all(np.abs(rec - w) <= 0.08 * np.abs(w) + 1e-3)


def test_fp4_emits_packed_weight_and_group_scales():
    """:1055 PerGroupFp4_EmitsNibblePackedWeightAndPerGroupScales (+ the nibble values, which the
    reference test does not check)."""
    N, K = 32, 256
    lin = Linear("l", LinearConfig(K, N).withBias(False), "cuda:0", PerGroupFp4(128))
    lin.build((1, K)); blob, bits = _blob(N, K); lin.loadParameter("weight", blob)
    t = lin.saveFlatTensors("l")
    assert t["l.weight"].shape == (N, K // 2) and t["l.weight"].dtype == torch.uint8
    s = t["l.weight_scale"]
    assert s.shape == (N, K // 128) and bool(torch.isfinite(s).all()) and bool((s > 0).all())
    qo, so = O.quantize_fp4_per_group(bits, 128)
    np.testing.assert_array_equal(t["l.weight"].numpy(), qo)
    np.testing.assert_array_equal(s.numpy(), so)


@pytest.mark.parametrize("policy,M", [(PerChannelFp8(), 1), (PerGroupFp4(128), 16)], ids=["fp8", "fp4_m16"])
def test_prequantized_reload_is_byte_identical_and_forward_equal(policy, M):
    """:1145 PreQuantizedArtifactLoadsBackWithoutRequantizing, :1304 PreQuantizedFp4Reload_Forward..."""
    N, K = 64, 256
    cfg = LinearConfig(K, N).withBias(False)
    a = Linear("a", cfg, "cuda:0", policy); a.build((M, K))
    blob, _ = _blob(N, K); a.loadParameter("weight", blob)
    saved = a.saveFlatTensors("a")
    b = Linear("b", cfg, "cuda:0", policy); b.build((M, K))
    b.loadParameter("weight", TensorBlob(policy.kStorageDtype, tuple(saved["a.weight"].shape), saved["a.weight"]))
    b.loadParameter("weight_scale", TensorBlob("FP32", tuple(saved["a.weight_scale"].shape), saved["a.weight_scale"]))
    again = b.saveFlatTensors("a")
    assert torch.equal(saved["a.weight"], again["a.weight"])
    assert torch.equal(saved["a.weight_scale"], again["a.weight_scale"])
    x = _spread((M, K))
    assert torch.equal(a.forward(x), b.forward(x))


def test_error_conventions():
    cfg = LinearConfig(K_IN, K_OUT).withBias(False)
    lin = Linear("l", cfg, "cuda:0", PerChannelFp8())
    with pytest.raises(_lib.MilaB200Error):
        lin.forward(_spread((1, K_IN)))                              # not built: runtime_error (:163)
    lin.build((1, K_IN))
    with pytest.raises(_lib.InvalidArgument):
        lin.forward(_spread((1, K_IN + 64)))                         # feature mismatch (:979)
    blob, _ = _blob(K_OUT + 1, K_IN)
    with pytest.raises(_lib.InvalidArgument):
        lin.loadParameter("weight", blob)                            # shape mismatch (Quantize.ixx:67-73)
    with pytest.raises(_lib.InvalidArgument):
        lin.loadParameter("gamma", blob)                             # unknown parameter (:595)
    with pytest.raises(_lib.LogicError):
        lin.backward()                                               # :225-228
    with pytest.raises(_lib.InvalidArgument):
        Linear("g", LinearConfig(192, 32).withBias(False), "cuda:0", PerGroupFp4(128)).build((1, 192))


def test_runtime_shape_differs_from_build_shape_and_shared_output():
    """Linear.ixx:176-187 view return; :682-690 shared output slot wider than the result."""
    N, K = 48, 128
    lin = Linear("l", LinearConfig(K, N).withBias(True), "cuda:0", PerGroupFp4(64))
    slot = torch.full((4 * 16 * N,), 7.0, dtype=torch.bfloat16, device="cuda")
    lin.installSharedOutput(slot)
    lin.build((2, 4, K)); blob, bits = _blob(N, K); lin.loadParameter("weight", blob)
    lin.loadParameter("bias", TensorBlob("BF16", (N,), G.bf16_tensor(O.f32_to_bf16_bits(O.ref_bias_value(np.arange(N))), "cpu")))
    y = lin.forward(_spread((2, 4, K)))
    assert y.shape == (2, 4, N) and y.data_ptr() == slot.data_ptr()
    assert bool((slot[8 * N:] == 7.0).all())                         # exactly M*N elements written
    y1 = lin.forward(_spread((1, 1, K)))
    assert y1.shape == (1, 1, N)
    assert lin.getRequiredMemory() == N * K // 2 + 4 * N * (K // 64) + 2 * N


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128), PerGroupFp4(64)], ids=["fp8", "fp4g128", "fp4g64"])
def test_artifact_file_round_trip_without_requantizing(policy, tmp_path):
    """Linear -> flat safetensors artifact -> Linear.  The reload copies packed bytes and scales as
    they are (Linear.ixx:543-574: storage-dtype blob = raw copy, then `weight_scale`), so the tensors are byte-equal,
    forwards are EXPECT_EQ-equal, and the file bytes equal this library's own quantizer output (= Mila's)."""
    from mila_b200 import artifact as A
    N, K, M = 64, 256, 3
    cfg_b, cfg_n = LinearConfig(K, N).withBias(True), LinearConfig(K, 2 * N).withBias(False)
    a = Linear("a", cfg_b, "cuda:0", policy); a.build((M, K))
    c = Linear("c", cfg_n, "cuda:0", policy); c.build((M, K))
    blob_a, bits_a = _blob(N, K); a.loadParameter("weight", blob_a)
    a.loadParameter("bias", TensorBlob("BF16", (N,), G.bf16_tensor(O.f32_to_bf16_bits(O.ref_bias_value(np.arange(N))), "cpu")))
    blob_c, _ = _blob(2 * N, K); c.loadParameter("weight", blob_c)
    path = tmp_path / "tiny.safetensors"
    A.saveLinearArtifact(path, {"tf_layer_0.qkv_proj": a, "tf_layer_0.fc_gate_up": c}, policy, '{"architecture":"llama"}')

    with A.ArtifactReader(path) as r:
        assert r.getWeightQuantization() == policy.tag
        m = r.getTensorMetadata("tf_layer_0.qkv_proj.weight")
        assert m.dtype == policy.kStorageDtype and m.shape == tuple(a.weight_.shape)
        assert r.getTensorMetadata("tf_layer_0.qkv_proj.weight_scale").dtype == "FP32"
        assert r.getTensorMetadata("tf_layer_0.qkv_proj.bias").dtype == "BF16"
        assert not r.hasTensor("tf_layer_0.fc_gate_up.bias")
        # the bytes in the file are the oracle's (= the reference kernels') packing of the BF16 source
        if isinstance(policy, PerChannelFp8):
            qo, so = O.quantize_fp8_per_channel(bits_a)
        else:
            qo, so = O.quantize_fp4_per_group(bits_a, policy.kQuantizationGroupSize)
        np.testing.assert_array_equal(r.readTensorBlob("tf_layer_0.qkv_proj.weight").data.numpy(), qo)
        np.testing.assert_array_equal(r.readTensorBlob("tf_layer_0.qkv_proj.weight_scale").data.numpy(), so)

        a2 = Linear("a2", cfg_b, "cuda:0", policy); a2.build((M, K))
        c2 = Linear("c2", cfg_n, "cuda:0", policy); c2.build((M, K))
        _lib.reset_launch_count()
        A.loadLinearFromArtifact(r, "tf_layer_0.qkv_proj", a2)
        A.loadLinearFromArtifact(r, "tf_layer_0.fc_gate_up", c2)
        assert _lib.launch_count() == 0                              # no quantizer ran: raw copies only

        wrong = PerGroupFp4(64) if policy != PerGroupFp4(64) else PerGroupFp4(128)
        other = Linear("w", cfg_n, "cuda:0", wrong); other.build((M, K))
        with pytest.raises(_lib.MilaB200Error):                       # GemmaModel.ixx:612-633
            A.loadLinearFromArtifact(r, "tf_layer_0.fc_gate_up", other)
        with pytest.raises(_lib.MilaB200Error):
            A.loadLinearFromArtifact(r, "tf_layer_9.missing", c2)

    for p_, q_ in ((a, a2), (c, c2)):
        assert torch.equal(p_.weight_, q_.weight_) and torch.equal(p_.weight_scales_, q_.weight_scales_)
    assert torch.equal(a.bias_, a2.bias_)
    x = _spread((M, K))
    assert torch.equal(a.forward(x), a2.forward(x)) and torch.equal(c.forward(x), c2.forward(x))


def test_artifact_with_bf16_source_quantizes_on_load(tmp_path):
    """An artifact that declares no scheme ("none") carries compute-precision weights: any policy may quantize it on
    load (GemmaModel.ixx:616-618), and the result equals quantizing the same blob directly."""
    from mila_b200 import artifact as A
    N, K = 32, 256
    blob, bits = _blob(N, K)
    path = tmp_path / "bf16.safetensors"
    w = A.SafeTensorsWriter(path)
    w.setMetadata(A.kMilaQuantizationMetadataKey, "none")
    w.declareTensor("l.weight", "BF16", (N, K))
    w.beginData(); w.writeTensorData("l.weight", blob.data); w.close()
    for policy in (PerChannelFp8(), PerGroupFp4(128)):
        direct = Linear("d", LinearConfig(K, N).withBias(False), "cuda:0", policy); direct.build((1, K))
        direct.loadParameter("weight", blob)
        via = Linear("v", LinearConfig(K, N).withBias(False), "cuda:0", policy); via.build((1, K))
        with A.ArtifactReader(path) as r:
            assert r.getWeightQuantization() == ""
            A.loadLinearFromArtifact(r, "l", via)                    # pageable mmap view as the quantizer's source
        assert torch.equal(direct.weight_, via.weight_) and torch.equal(direct.weight_scales_, via.weight_scales_)
