"""Pre-quantized artifact I/O for the quantized Linear path.

Mila distributes models as ONE flat safetensors file whose Linear tensors are already packed:
`<prefix>.weight` (F8_E4M3 `[N, K]` or U8 `[N, K/2]`), `<prefix>.weight_scale` (F32 `[N]` or `[N, K/g]`),
`<prefix>.bias` (BF16 `[N]`), with the policy name in `__metadata__["mila_quantization"]`
(Serialization/SafeTensors.ixx, Components/Linear/Linear.ixx:370-400 save, :529-600 load,
Core/LanguageModel.ixx:117-147 savePretrained, Serialization/PretrainedReader.ixx:980-1125 read).
Because this library's quantizers are bit-exact with Mila's, such artifacts are interchangeable in both
directions; this module is the host side that reads and writes them WITHOUT touching the BF16 source
(4x the FP4 bytes over PCIe, one stream sync per tensor — the load-time cost §8a4/a5 names).

Mirrored behaviour (so that the tests read like Mila/Tests/Dnn/Serialization/SafeTensors.Cpu.cpp):
  * two-phase writer: declare every tensor, beginData(), then write bodies in declaration order; duplicate names,
    late declarations, out-of-order or wrong-size bodies and an incomplete close() raise (runtime_error);
  * header = compact JSON, keys sorted (nlohmann::json objects are std::map), space-padded so that the data
    region starts 8-byte aligned; tensors are laid back to back in declaration order (no per-tensor padding):
    a file written here is byte-identical to SafeTensorsWriter's for the same declarations;
  * reader: u64 header length in (0, 128 MiB], JSON object, every entry needs dtype / shape / data_offsets,
    rank <= 8, offsets ordered and inside the file; "none" quantization reads as absent ("");
  * a pre-quantized artifact is refused when its declared policy differs from the requested one
    (Models/GemmaModel.ixx:612-633) — FP4 g=128 and g=64 are both U8 and only the name tells them apart.
"""
from __future__ import annotations

import json
import mmap
import os
import struct
import warnings
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

from ._lib import InvalidArgument, MilaB200Error
from .linear import Linear, PerChannelFp8, PerGroupFp4, TensorBlob

kMilaConfigMetadataKey = "mila_config"                   # SafeTensors.ixx:36
kMilaQuantizationMetadataKey = "mila_quantization"       # SafeTensors.ixx:49
MAX_SAFETENSORS_HEADER_BYTES = 128 * 1024 * 1024         # PretrainedReader.ixx:503
MAX_RANK = 8                                             # TensorShape::MaxRank

# Mila TensorDataType name -> (safetensors spelling, storage bytes per element); SafeTensors.ixx:59-113.
# Packed 4-bit weights travel as U8 with their physical (halved) column count.
_DTYPES = {
    "FP32": ("F32", 4), "FP16": ("F16", 2), "BF16": ("BF16", 2), "FP8_E4M3": ("F8_E4M3", 1), "FP8_E5M2": ("F8_E5M2", 1),
    "INT8": ("I8", 1), "INT16": ("I16", 2), "INT32": ("I32", 4), "UINT8": ("U8", 1), "UINT16": ("U16", 2), "UINT32": ("U32", 4),
}
_FROM_ST = {st: name for name, (st, _) in _DTYPES.items()}
_TORCH_VIEW = {"FP32": torch.float32, "FP16": torch.float16, "BF16": torch.bfloat16, "FP8_E4M3": torch.uint8,
               "FP8_E5M2": torch.uint8, "INT8": torch.int8, "INT16": torch.int16, "INT32": torch.int32,
               "UINT8": torch.uint8, "UINT16": torch.int16, "UINT32": torch.int32}


def toSafeTensorsDataTypeName(dtype: str) -> str:
    if dtype not in _DTYPES:
        raise MilaB200Error(f"safetensors: dtype {dtype} has no container spelling")
    return _DTYPES[dtype][0]


def fromSafeTensorsDataTypeName(name: str) -> str:
    if name not in _FROM_ST:
        raise MilaB200Error(f"safetensors: unsupported dtype '{name}'")
    return _FROM_ST[name]


def storageBytesPerElement(dtype: str) -> int:
    if dtype not in _DTYPES:
        raise MilaB200Error(f"safetensors: no storage width for dtype {dtype}")
    return _DTYPES[dtype][1]


def weightQuantizationName(policy) -> str:
    """Core/LanguageModelConfig.ixx:104-114.  The reference ships exactly two quantized schemes; FP4 at group 64
    exists as a policy (Policies.ixx:104-113) but has no artifact name, so it gets the obvious spelling here."""
    if policy is None:
        return "none"
    if isinstance(policy, (PerChannelFp8, PerGroupFp4)):
        return policy.tag
    raise InvalidArgument(f"unknown weight quantization policy {policy!r}")


DECLARE, WRITE = "declare", "write"                      # TensorSavePass (SafeTensors.ixx:138-143)


class SafeTensorsWriter:
    """SafeTensors.ixx:146-: declareTensor()* -> beginData() -> writeTensorData()* (same order) -> close()."""

    def __init__(self, filepath: str | os.PathLike):
        self._path = Path(filepath)
        if str(self._path.parent):
            self._path.parent.mkdir(parents=True, exist_ok=True)
        try:
            self._file = open(self._path, "wb")
        except OSError as e:
            raise MilaB200Error(f"Cannot open safetensors file for writing: {self._path}") from e
        self._entries: list[dict] = []
        self._metadata: dict[str, str] = {}
        self._next_offset = 0
        self._header_written = False
        self._next_write_index = 0

    def __enter__(self): return self

    def __exit__(self, exc_type, *_):
        if exc_type is None:
            self.close()
        elif self._file is not None:
            self._file.close(); self._file = None

    def declareTensor(self, name: str, dtype: str, shape) -> None:
        if self._header_written:
            raise MilaB200Error(f"SafeTensorsWriter: cannot declare '{name}' after beginData()")
        if any(e["name"] == name for e in self._entries):
            raise MilaB200Error(f"SafeTensorsWriter: duplicate tensor name '{name}'")
        count = 1
        for d in shape:
            if int(d) < 0:
                raise MilaB200Error(f"SafeTensorsWriter: negative extent in shape of '{name}'")
            count *= int(d)
        nbytes = count * storageBytesPerElement(dtype)
        self._entries.append({"name": name, "dtype": dtype, "shape": [int(d) for d in shape],
                              "begin": self._next_offset, "nbytes": nbytes})
        self._next_offset += nbytes

    def setMetadata(self, key: str, value: str) -> None:
        if self._header_written:
            raise MilaB200Error(f"SafeTensorsWriter: cannot set metadata '{key}' after beginData()")
        self._metadata[key] = value

    def beginData(self) -> None:
        if self._header_written:
            raise MilaB200Error("SafeTensorsWriter: beginData() called twice")
        header: dict = {}
        if self._metadata:
            header["__metadata__"] = dict(self._metadata)
        for e in self._entries:
            header[e["name"]] = {"dtype": toSafeTensorsDataTypeName(e["dtype"]), "shape": e["shape"],
                                 "data_offsets": [e["begin"], e["begin"] + e["nbytes"]]}
        # nlohmann::json::dump(): compact, object keys in std::map (byte) order, UTF-8 passed through
        text = json.dumps(header, separators=(",", ":"), sort_keys=True, ensure_ascii=False).encode("utf-8")
        while (8 + len(text)) % 8 != 0:                   # the data region must start 8-byte aligned
            text += b" "
        self._file.write(struct.pack("<Q", len(text)))
        self._file.write(text)
        self._header_written = True
        self._next_write_index = 0

    def writeTensorData(self, name: str, data) -> None:
        """`data`: bytes-like, numpy array or CPU torch tensor holding exactly the declared bytes."""
        if not self._header_written:
            raise MilaB200Error(f"SafeTensorsWriter: writeTensorData('{name}') before beginData()")
        if self._next_write_index >= len(self._entries):
            raise MilaB200Error(f"SafeTensorsWriter: unexpected extra tensor '{name}'")
        e = self._entries[self._next_write_index]
        if e["name"] != name:
            raise MilaB200Error(f"SafeTensorsWriter: out of order write; expected '{e['name']}', got '{name}'")
        buf = _as_bytes(data)
        if len(buf) != e["nbytes"]:
            raise MilaB200Error(f"SafeTensorsWriter: '{name}' declared {e['nbytes']} bytes, got {len(buf)}")
        if len(buf):
            self._file.write(buf)
        self._next_write_index += 1

    def close(self) -> None:
        if self._file is None:
            return
        incomplete = self._header_written and self._next_write_index != len(self._entries)
        missing = self._entries[self._next_write_index]["name"] if incomplete else ""
        self._file.close(); self._file = None
        if incomplete:
            raise MilaB200Error(f"SafeTensorsWriter: closed with {self._next_write_index} of {len(self._entries)} "
                                f"tensors written; '{missing}' missing")

    def getTensorCount(self) -> int:
        return len(self._entries)


def _as_bytes(data) -> memoryview:
    if isinstance(data, torch.Tensor):
        if data.device.type != "cpu":
            raise InvalidArgument("writeTensorData: tensor must be on the host")
        data = data.contiguous().view(torch.uint8).numpy()
    if isinstance(data, np.ndarray):
        return memoryview(np.ascontiguousarray(data).view(np.uint8).reshape(-1))
    return memoryview(data).cast("B")


@dataclass
class TensorBlobMetadata:                                 # Serialization/TensorBlob: name, dtype, shape, byte range
    name: str
    dtype: str
    shape: tuple
    offset: int
    nbytes: int


class ArtifactReader:
    """The safetensors branch of PretrainedModelReader (PretrainedReader.ixx:980-1125): header parse with the
    reference's validation, a tensor index, and zero-copy blobs over one read-only mapping of the file."""

    def __init__(self, filepath: str | os.PathLike):
        self._path = Path(filepath)
        try:
            self._fd = open(self._path, "rb")
        except OSError as e:
            raise MilaB200Error(f"Cannot open pretrained model file: {self._path}") from e
        file_size = os.fstat(self._fd.fileno()).st_size
        head = self._fd.read(8)
        if len(head) != 8:
            self._fd.close()
            raise MilaB200Error(f"Failed reading safetensors header length from {self._path}")
        (header_length,) = struct.unpack("<Q", head)
        if header_length == 0 or header_length > MAX_SAFETENSORS_HEADER_BYTES or 8 + header_length > file_size:
            self._fd.close()
            raise MilaB200Error(f"{self._path} is not a safetensors container: header length {header_length} is out of range")
        text = self._fd.read(header_length)
        try:
            header = json.loads(text.decode("utf-8"))
        except (UnicodeDecodeError, json.JSONDecodeError):
            header = None
        if not isinstance(header, dict):
            self._fd.close()
            raise MilaB200Error("Malformed safetensors header: not a JSON object")
        data_start = 8 + header_length
        self._index: dict[str, TensorBlobMetadata] = {}
        self._weight_quantization = ""
        self._mila_config: str | None = None
        try:
            for name, record in header.items():
                if name == "__metadata__":
                    self._read_metadata(record)
                    continue
                if not isinstance(record, dict) or not all(k in record for k in ("dtype", "shape", "data_offsets")):
                    raise MilaB200Error(f"Malformed safetensors entry for '{name}'")
                dims, offsets = record["shape"], record["data_offsets"]
                if not isinstance(dims, list) or len(dims) > MAX_RANK:
                    raise MilaB200Error(f"Invalid tensor dimensionality for '{name}'")
                if not isinstance(offsets, list) or len(offsets) != 2:
                    raise MilaB200Error(f"Invalid data_offsets for '{name}'")
                begin, end = int(offsets[0]), int(offsets[1])
                if end < begin:
                    raise MilaB200Error(f"Inverted data_offsets for '{name}'")
                meta = TensorBlobMetadata(name, fromSafeTensorsDataTypeName(record["dtype"]),
                                          tuple(int(d) for d in dims), data_start + begin, end - begin)
                if meta.offset + meta.nbytes > file_size:
                    raise MilaB200Error(f"Tensor '{name}' extends past end of file")
                count = 1
                for d in meta.shape:
                    if d < 0:
                        raise MilaB200Error(f"Negative dimension in shape of '{name}'")
                    count *= d
                if meta.dtype in _TORCH_VIEW and meta.nbytes != count * storageBytesPerElement(meta.dtype):
                    raise MilaB200Error(f"Tensor '{name}': data_offsets span {meta.nbytes} bytes but shape {list(meta.shape)} "
                                        f"of {meta.dtype} needs {count * storageBytesPerElement(meta.dtype)}")
                self._index[name] = meta
            if not self._index:
                raise MilaB200Error("safetensors file declares no tensors")
        except Exception:
            self._fd.close()
            raise
        self._map = mmap.mmap(self._fd.fileno(), 0, access=mmap.ACCESS_READ) if file_size else None

    def _read_metadata(self, metadata) -> None:
        if not isinstance(metadata, dict):
            return
        q = metadata.get(kMilaQuantizationMetadataKey)
        if isinstance(q, str):
            self._weight_quantization = "" if q == "none" else q          # one test for "quantize on load"
        if kMilaConfigMetadataKey in metadata:
            cfg = metadata[kMilaConfigMetadataKey]
            if not isinstance(cfg, str):
                raise MilaB200Error(f"safetensors __metadata__['{kMilaConfigMetadataKey}'] is not a string")
            self._mila_config = cfg

    # -- PretrainedModelReader surface used by the Linear path
    def getTensorNames(self) -> list[str]: return list(self._index)
    def hasTensor(self, name: str) -> bool: return name in self._index
    def getWeightQuantization(self) -> str: return self._weight_quantization
    def getMilaConfigJSON(self) -> str | None: return self._mila_config

    def getTensorMetadata(self, name: str) -> TensorBlobMetadata:
        if name not in self._index:
            raise MilaB200Error(f"Tensor '{name}' not found in {self._path}")
        return self._index[name]

    def readTensorBlob(self, name: str) -> TensorBlob:
        """Zero-copy view of the tensor's bytes in the file mapping (pageable host memory — the launchers accept
        that, SURVEY.md §8b 'Ownership'); dtype tag and shape as recorded."""
        m = self.getTensorMetadata(name)
        raw = np.frombuffer(self._map, dtype=np.uint8, count=m.nbytes, offset=m.offset)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)   # read-only mapping: consumers only read
            t = torch.from_numpy(raw)
        want = _TORCH_VIEW[m.dtype]
        if want != torch.uint8:
            if m.offset % torch.empty((), dtype=want).element_size() != 0:
                t = t.clone()                              # unaligned for its element type: one host copy
            t = t.view(want)
        return TensorBlob(m.dtype, m.shape, t.reshape(m.shape) if m.nbytes else t)

    def close(self) -> None:
        if getattr(self, "_map", None) is not None:
            self._map = None                               # views keep the mapping alive; drop our handle
        if self._fd is not None:
            self._fd.close(); self._fd = None

    def __enter__(self): return self
    def __exit__(self, *_): self.close()


# ---- Linear glue ---------------------------------------------------------------------------------------------------

def saveLinearFlatTensors(linear: Linear, writer: SafeTensorsWriter, prefix: str, save_pass: str) -> None:
    """Linear::saveFlatTensors (Linear.ixx:370-400) through the two-phase writer: ONE ordered body for both
    passes.  A borrowed (tied) weight is not emitted."""
    policy = linear._policy
    tensors = linear.saveFlatTensors(prefix) if save_pass == WRITE else None
    names = []
    if not linear._weight_installed:
        names.append((prefix + ".weight", policy.kStorageDtype, tuple(linear.weight_.shape)))
        names.append((prefix + ".weight_scale", "FP32", tuple(linear.weight_scales_.shape)))
    if linear.bias_ is not None:
        names.append((prefix + ".bias", "BF16", tuple(linear.bias_.shape)))
    for name, dtype, shape in names:
        if save_pass == DECLARE:
            writer.declareTensor(name, dtype, shape)
        else:
            writer.writeTensorData(name, tensors[name])


def saveLinearArtifact(path, linears: dict[str, Linear], policy, mila_config_json: str | None = None) -> None:
    """LanguageModel::savePretrained (Core/LanguageModel.ixx:117-147) restricted to Linear components:
    metadata, declare pass, beginData, write pass, close.  The file is built next to its destination and renamed on
    success, so a failed save never leaves a truncated artifact at `path`."""
    path = Path(path)
    tmp = path.with_name(path.name + ".partial")
    writer = SafeTensorsWriter(tmp)
    try:
        if mila_config_json is not None:
            writer.setMetadata(kMilaConfigMetadataKey, mila_config_json)
        writer.setMetadata(kMilaQuantizationMetadataKey, weightQuantizationName(policy))
        for prefix, lin in linears.items():
            saveLinearFlatTensors(lin, writer, prefix, DECLARE)
        writer.beginData()
        for prefix, lin in linears.items():
            lin.synchronize()
            saveLinearFlatTensors(lin, writer, prefix, WRITE)
        writer.close()
    except Exception:
        if writer._file is not None:
            writer._file.close(); writer._file = None
        tmp.unlink(missing_ok=True)
        raise
    os.replace(tmp, path)


def loadLinearFromArtifact(reader: ArtifactReader, prefix: str, linear: Linear) -> None:
    """The per-component slice of fromPretrained (GemmaModel.ixx:612-633 policy check, CompositeComponent load
    loop): refuse a mismatched scheme, then route `weight`, `weight_scale`, `bias` through loadParameter — packed
    bytes are copied as they are, a BF16 source (no declared scheme) is quantized on load."""
    declared = reader.getWeightQuantization()
    requested = weightQuantizationName(linear._policy)
    if declared and declared != requested:
        raise MilaB200Error(f"artifact '{reader._path}' is pre-quantized as '{declared}' but this load requested '{requested}'")
    if not linear._weight_installed:
        if not reader.hasTensor(prefix + ".weight"):
            raise MilaB200Error(f"Tensor '{prefix}.weight' not found in {reader._path}")
        linear.loadParameter("weight", reader.readTensorBlob(prefix + ".weight"))
        if reader.hasTensor(prefix + ".weight_scale"):
            linear.loadParameter("weight_scale", reader.readTensorBlob(prefix + ".weight_scale"))
        elif declared:
            raise MilaB200Error(f"pre-quantized artifact carries no '{prefix}.weight_scale'")
    if linear.hasBias() and reader.hasTensor(prefix + ".bias"):
        linear.loadParameter("bias", reader.readTensorBlob(prefix + ".bias"))
    linear.synchronize()                                   # the mapping may go away once the caller closes the reader


# ---- tensor-parallel shards straight from the artifact ----------------------------------------------

def readLinearShard(reader: ArtifactReader, prefix: str, policy, world: int, rank: int, split: str):
    """This rank's shard of a pre-quantized Linear, cut from the file mapping: `split` = "column" (rows
    [r N/p, (r+1) N/p): QKV / gate / up) or "row" (a K slice holding whole FP4 groups: o_proj / down).  Returns host
    tensors (weight_shard, scale_shard, bias_or_None) ready for an H2D copy — no rank ever materialises the unsharded
    matrix.  The slicing rules are tp.column_shard / tp.row_shard's: shards are bit-exact slices of the unsharded
    quantisation; an FP8 row-parallel shard keeps the FULL per-channel scale vector (the scale is the absmax of the whole
    row).  Every rank of a row-parallel layer gets the FULL bias: TpGroup.rowparallel_forward adds it exactly once whatever
    route it takes (the fused decode all-reduce adds it after the cross-rank sum on every rank; the NCCL route masks it to
    rank 0 before the sum), so the outputs are identical on all ranks."""
    from .tp import column_shard, row_shard
    declared = reader.getWeightQuantization()
    requested = weightQuantizationName(policy)
    if declared != requested:
        raise MilaB200Error(f"artifact '{reader._path}' is pre-quantized as '{declared or 'none'}' but this load requested "
                            f"'{requested}': shards can only be cut from packed tensors of the same scheme")
    w = reader.readTensorBlob(prefix + ".weight")
    sc = reader.readTensorBlob(prefix + ".weight_scale")
    if w.dtype != policy.kStorageDtype or sc.dtype != "FP32":
        raise InvalidArgument(f"'{prefix}': weight dtype {w.dtype} / scale dtype {sc.dtype} do not match {requested}")
    wt, st = w.data.view(torch.uint8), sc.data
    if split == "column":
        ws, ss = column_shard(wt, st, world, rank)
    elif split == "row":
        ws, ss = row_shard(wt, st, policy, world, rank)
        ss = ss.clone() if ss.data_ptr() == st.data_ptr() else ss     # never hand out a view of the read-only mapping
    else:
        raise InvalidArgument(f"split must be 'column' or 'row', got {split!r}")
    bias = None
    if reader.hasTensor(prefix + ".bias"):
        b = reader.readTensorBlob(prefix + ".bias").data
        if split == "column":
            from .tp import shard_bounds
            bias = b[shard_bounds(b.shape[0], world, rank)].clone()
        else:
            bias = b.clone()
    return ws, ss, bias
