"""Flat-safetensors artifact container (mila_b200/artifact.py) — CPU tests mirroring
Mila/Tests/Dnn/Serialization/SafeTensors.Cpu.cpp (:133 RoundTripsTensorsOfMixedDataTypes, :320/:340 declared
quantization, :384-:456 writer misuse, :463/:478 malformed files), plus the byte layout a foreign reader relies on."""
import json
import struct

import numpy as np
import pytest
import torch

from mila_b200 import artifact as A
from mila_b200._lib import MilaB200Error
from mila_b200.linear import PerChannelFp8, PerGroupFp4


def _write_three(path):
    weights = np.arange(6, dtype=np.float32) * 0.5 - 1.0
    packed = np.array([0x12, 0xAB, 0xFF, 0x00, 0x7E], dtype=np.uint8)
    counts = np.array([3, -7], dtype=np.int32)
    w = A.SafeTensorsWriter(path)
    w.declareTensor("layer.weight", "FP32", (2, 3))
    w.declareTensor("layer.packed", "UINT8", (5,))
    w.declareTensor("layer.counts", "INT32", (2,))
    assert w.getTensorCount() == 3
    w.beginData()
    w.writeTensorData("layer.weight", weights)
    w.writeTensorData("layer.packed", packed)
    w.writeTensorData("layer.counts", counts)
    w.close()
    return weights, packed, counts


def test_round_trips_tensors_of_mixed_data_types(tmp_path):
    f = tmp_path / "m.safetensors"
    weights, packed, counts = _write_three(f)
    with A.ArtifactReader(f) as r:
        assert sorted(r.getTensorNames()) == ["layer.counts", "layer.packed", "layer.weight"]
        m = r.getTensorMetadata("layer.weight")
        assert m.dtype == "FP32" and m.shape == (2, 3) and m.nbytes == 24
        np.testing.assert_array_equal(r.readTensorBlob("layer.weight").data.numpy(), weights.reshape(2, 3))
        b = r.readTensorBlob("layer.packed")
        assert b.dtype == "UINT8" and b.data.numpy().tobytes() == packed.tobytes()
        c = r.readTensorBlob("layer.counts")
        assert c.dtype == "INT32" and c.data.numpy().tolist() == [3, -7]
        assert r.getWeightQuantization() == "" and r.getMilaConfigJSON() is None
        with pytest.raises(MilaB200Error):
            r.getTensorMetadata("nope")


def test_file_layout_is_what_a_foreign_safetensors_reader_expects(tmp_path):
    """u64 little-endian header length, compact JSON with sorted keys padded with spaces so that the data region
    is 8-byte aligned, tensors back to back in declaration order (SafeTensors.ixx:243-283)."""
    f = tmp_path / "m.safetensors"
    weights, packed, counts = _write_three(f)
    raw = f.read_bytes()
    (hl,) = struct.unpack("<Q", raw[:8])
    assert (8 + hl) % 8 == 0
    text = raw[8:8 + hl].decode()
    assert text.rstrip(" ") == json.dumps(json.loads(text), separators=(",", ":"), sort_keys=True)
    hdr = json.loads(text)
    assert hdr["layer.weight"] == {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 24]}
    assert hdr["layer.packed"] == {"dtype": "U8", "shape": [5], "data_offsets": [24, 29]}
    assert hdr["layer.counts"] == {"dtype": "I32", "shape": [2], "data_offsets": [29, 37]}    # no per-tensor padding
    assert raw[8 + hl:] == weights.tobytes() + packed.tobytes() + counts.tobytes()
    # the torch-free reference parser of the format agrees
    try:
        from safetensors import safe_open
    except ImportError:
        return
    with safe_open(str(f), framework="np") as sf:
        np.testing.assert_array_equal(sf.get_tensor("layer.weight"), weights.reshape(2, 3))
        np.testing.assert_array_equal(sf.get_tensor("layer.counts"), counts)


@pytest.mark.parametrize("policy,name", [(PerGroupFp4(128), "per_group_fp4_128"), (PerChannelFp8(), "per_channel_fp8_e4m3"),
                                         (PerGroupFp4(64), "per_group_fp4_64"), (None, "none")])
def test_surfaces_the_declared_weight_quantization(tmp_path, policy, name):
    """:320 SurfacesTheDeclaredWeightQuantization, :340 TreatsAnUnquantizedDeclarationAsAbsent; the names are
    LanguageModelConfig.ixx:104-114's."""
    assert A.weightQuantizationName(policy) == name
    f = tmp_path / "q.safetensors"
    w = A.SafeTensorsWriter(f)
    w.setMetadata(A.kMilaQuantizationMetadataKey, name)
    w.setMetadata(A.kMilaConfigMetadataKey, json.dumps({"architecture": "llama", "num_layers": 2}))
    w.declareTensor("t", "BF16", (4,))
    w.beginData()
    w.writeTensorData("t", torch.ones(4, dtype=torch.bfloat16))
    w.close()
    with A.ArtifactReader(f) as r:
        assert r.getWeightQuantization() == ("" if name == "none" else name)
        assert json.loads(r.getMilaConfigJSON())["architecture"] == "llama"
        assert r.readTensorBlob("t").data.dtype == torch.bfloat16
    hdr = json.loads(f.read_bytes()[8:8 + struct.unpack("<Q", f.read_bytes()[:8])[0]])
    assert list(hdr)[0] == "__metadata__"                 # '_' sorts before lower-case tensor names


def test_fp8_dtype_spelling_and_packed_shapes(tmp_path):
    """FP8 weights are F8_E4M3 [N, K]; packed FP4 is U8 with the physical halved column count (SafeTensors.ixx:52-57)."""
    f = tmp_path / "w.safetensors"
    w = A.SafeTensorsWriter(f)
    w.declareTensor("a.weight", "FP8_E4M3", (4, 16))
    w.declareTensor("a.weight_scale", "FP32", (4,))
    w.declareTensor("b.weight", "UINT8", (4, 8))
    w.declareTensor("b.weight_scale", "FP32", (4, 1))
    w.beginData()
    w.writeTensorData("a.weight", np.arange(64, dtype=np.uint8))
    w.writeTensorData("a.weight_scale", np.ones(4, np.float32))
    w.writeTensorData("b.weight", np.arange(32, dtype=np.uint8))
    w.writeTensorData("b.weight_scale", np.ones((4, 1), np.float32))
    w.close()
    with A.ArtifactReader(f) as r:
        assert r.getTensorMetadata("a.weight").dtype == "FP8_E4M3" and r.getTensorMetadata("a.weight").shape == (4, 16)
        assert r.readTensorBlob("a.weight").data.dtype == torch.uint8
        assert r.getTensorMetadata("b.weight").dtype == "UINT8" and r.getTensorMetadata("b.weight").nbytes == 32
    assert b'"F8_E4M3"' in f.read_bytes()


def test_writer_rejects_misuse(tmp_path):
    """:384 out-of-order bodies, :402 size mismatch, :417 duplicates, :429 late declaration, :442 incomplete close."""
    w = A.SafeTensorsWriter(tmp_path / "a.safetensors")
    w.declareTensor("x", "FP32", (2,)); w.declareTensor("y", "FP32", (2,))
    with pytest.raises(MilaB200Error):
        w.declareTensor("x", "FP32", (2,))
    with pytest.raises(MilaB200Error):
        w.declareTensor("z", "FP32", (-1,))
    with pytest.raises(MilaB200Error):
        w.writeTensorData("x", np.zeros(2, np.float32))              # before beginData()
    w.beginData()
    with pytest.raises(MilaB200Error):
        w.beginData()
    with pytest.raises(MilaB200Error):
        w.declareTensor("late", "FP32", (1,))
    with pytest.raises(MilaB200Error):
        w.setMetadata("k", "v")
    with pytest.raises(MilaB200Error):
        w.writeTensorData("y", np.zeros(2, np.float32))              # out of order
    with pytest.raises(MilaB200Error):
        w.writeTensorData("x", np.zeros(3, np.float32))              # wrong size
    w.writeTensorData("x", np.zeros(2, np.float32))
    with pytest.raises(MilaB200Error):
        w.close()                                                    # 'y' never written
    w2 = A.SafeTensorsWriter(tmp_path / "b.safetensors")
    w2.declareTensor("x", "FP32", (1,)); w2.beginData(); w2.writeTensorData("x", np.zeros(1, np.float32))
    with pytest.raises(MilaB200Error):
        w2.writeTensorData("extra", np.zeros(1, np.float32))
    w2.close(); w2.close()                                           # idempotent
    with pytest.raises(MilaB200Error):
        A.toSafeTensorsDataTypeName("FP4_E2M1")                      # no container spelling: packed FP4 travels as U8


def _raw_file(path, header: dict | bytes, body=b""):
    text = header if isinstance(header, bytes) else json.dumps(header).encode()
    path.write_bytes(struct.pack("<Q", len(text)) + text + body)
    return path


def test_reader_rejects_malformed_files(tmp_path):
    """:463 RejectsAFileThatIsNeitherContainer, :478 RejectsATensorExtendingPastEndOfFile and the entry checks of
    PretrainedReader.ixx:1000-1075."""
    junk = tmp_path / "junk.bin"; junk.write_bytes(b"\xff" * 64)
    with pytest.raises(MilaB200Error):
        A.ArtifactReader(junk)
    short = tmp_path / "short.bin"; short.write_bytes(b"\x01\x02")
    with pytest.raises(MilaB200Error):
        A.ArtifactReader(short)
    with pytest.raises(MilaB200Error):
        A.ArtifactReader(tmp_path / "missing.safetensors")
    ok = {"dtype": "F32", "shape": [2], "data_offsets": [0, 8]}
    cases = {
        "past_end": ({"t": ok}, b"\0" * 4),
        "inverted": ({"t": dict(ok, data_offsets=[8, 0])}, b"\0" * 8),
        "no_dtype": ({"t": {"shape": [2], "data_offsets": [0, 8]}}, b"\0" * 8),
        "bad_offsets": ({"t": dict(ok, data_offsets=[0])}, b"\0" * 8),
        "rank": ({"t": dict(ok, shape=[1] * 9)}, b"\0" * 8),
        "dtype": ({"t": dict(ok, dtype="C64")}, b"\0" * 8),
        "empty": ({"__metadata__": {"mila_quantization": "none"}}, b""),
        "not_object": (b"[1, 2, 3]", b""),
        "not_json": (b"{not json", b""),
        "config_not_string": ({"__metadata__": {"mila_config": 5}, "t": ok}, b"\0" * 8),
    }
    for name, (hdr, body) in cases.items():
        with pytest.raises(MilaB200Error):
            A.ArtifactReader(_raw_file(tmp_path / f"{name}.safetensors", hdr, body))
    good = _raw_file(tmp_path / "good.safetensors", {"t": ok}, struct.pack("<2f", 1.5, -2.0))
    with A.ArtifactReader(good) as r:
        assert r.readTensorBlob("t").data.tolist() == [1.5, -2.0]


def test_unaligned_tensor_is_still_readable(tmp_path):
    """Tensors are laid back to back, so a 4-byte type can start at an odd offset (29 above): the blob is then a copy."""
    f = tmp_path / "m.safetensors"
    _, _, counts = _write_three(f)
    with A.ArtifactReader(f) as r:
        assert (r.getTensorMetadata("layer.counts").offset - r.getTensorMetadata("layer.weight").offset) == 29
        assert r.readTensorBlob("layer.counts").data.numpy().tolist() == counts.tolist()


# ---- tensor-parallel shards cut from the artifact (no GPU: pure slicing of the mapped file) -------------------------

def _packed_artifact(path, policy, N, K, with_bias):
    rng = np.random.default_rng(7)
    if isinstance(policy, PerChannelFp8):
        w = rng.integers(0, 256, (N, K), dtype=np.uint8); sc = rng.random(N, dtype=np.float32) + 0.5
        wdtype = "FP8_E4M3"
    else:
        g = policy.kQuantizationGroupSize
        w = rng.integers(0, 256, (N, K // 2), dtype=np.uint8); sc = rng.random((N, K // g), dtype=np.float32) + 0.5
        wdtype = "UINT8"
    bias = (rng.integers(0, 2 ** 15, N).astype(np.uint16)) if with_bias else None
    wr = A.SafeTensorsWriter(path)
    wr.setMetadata(A.kMilaQuantizationMetadataKey, A.weightQuantizationName(policy))
    wr.declareTensor("l.weight", wdtype, w.shape); wr.declareTensor("l.weight_scale", "FP32", sc.shape)
    if with_bias: wr.declareTensor("l.bias", "BF16", (N,))
    wr.beginData()
    wr.writeTensorData("l.weight", w); wr.writeTensorData("l.weight_scale", sc)
    if with_bias: wr.writeTensorData("l.bias", bias)
    wr.close()
    return w, sc, bias


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128), PerGroupFp4(64)], ids=["fp8", "fp4g128", "fp4g64"])
@pytest.mark.parametrize("world", [2, 8])
def test_shards_cut_from_the_artifact_tile_the_unsharded_tensors(tmp_path, policy, world):
    """column shards are row slices, row shards are K slices holding whole FP4 groups; FP8 row-parallel
    shards keep the full scale vector; every rank of a row-parallel layer gets the full bias (the forward adds it once).  Concatenating the
    shards gives back the file's tensors byte for byte."""
    N, K = 64, 2048
    f = tmp_path / "tp.safetensors"
    w, sc, bias = _packed_artifact(f, policy, N, K, with_bias=True)
    with A.ArtifactReader(f) as r:
        cols = [A.readLinearShard(r, "l", policy, world, k, "column") for k in range(world)]
        rows = [A.readLinearShard(r, "l", policy, world, k, "row") for k in range(world)]
    np.testing.assert_array_equal(np.concatenate([c[0].numpy() for c in cols], axis=0), w)
    np.testing.assert_array_equal(np.concatenate([c[1].numpy() for c in cols], axis=0), sc)
    np.testing.assert_array_equal(np.concatenate([c[2].view(torch.int16).numpy().view(np.uint16) for c in cols]), bias)
    np.testing.assert_array_equal(np.concatenate([c[0].numpy() for c in rows], axis=1), w)
    if isinstance(policy, PerChannelFp8):
        for c in rows:
            np.testing.assert_array_equal(c[1].numpy(), sc)            # whole-row absmax scale on every rank
    else:
        np.testing.assert_array_equal(np.concatenate([c[1].numpy() for c in rows], axis=1), sc)
        g = policy.kQuantizationGroupSize
        assert all(c[0].shape[1] * 2 % g == 0 for c in rows)           # whole groups per shard
    for c in rows:                                                       # full bias on every rank: rowparallel_forward adds it once
        np.testing.assert_array_equal(c[2].view(torch.int16).numpy().view(np.uint16), bias)
    assert all(t.is_contiguous() for c in cols + rows for t in c[:2])


def test_shard_loading_refuses_the_wrong_scheme_and_impossible_splits(tmp_path):
    f = tmp_path / "tp.safetensors"
    _packed_artifact(f, PerGroupFp4(128), 48, 1024, with_bias=False)
    with A.ArtifactReader(f) as r:
        with pytest.raises(MilaB200Error):
            A.readLinearShard(r, "l", PerGroupFp4(64), 2, 0, "column")      # both U8: only the name tells them apart
        with pytest.raises(MilaB200Error):
            A.readLinearShard(r, "l", PerChannelFp8(), 2, 0, "row")
        with pytest.raises(ValueError):
            A.readLinearShard(r, "l", PerGroupFp4(128), 5, 0, "column")     # 48 rows over 5 ranks
        with pytest.raises(ValueError):
            A.readLinearShard(r, "l", PerGroupFp4(128), 16, 0, "row")       # 1024 / 16 = 64 < one group of 128
        with pytest.raises(ValueError):
            A.readLinearShard(r, "l", PerGroupFp4(128), 2, 0, "diagonal")
        assert A.readLinearShard(r, "l", PerGroupFp4(128), 2, 1, "row")[2] is None
