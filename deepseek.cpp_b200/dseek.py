"""`.dseek` checkpoint container — reader/writer (pure numpy).

Byte-compatible with the reference's format: a checkpoint is a DIRECTORY whose every entry is a
safetensors-layout shard `u64 LE header_len | JSON header | raw tensor bytes`
(reader: /root/reference/src/codec.cpp:262-377, writer: convert.py:582-588).  `__metadata__` (flat
str->str map, convert.py:123-170) is taken from the first shard in sorted order; tensors are merged
from all shards.  K-quant tensors are stored as U8 with shape (rows, cols/256*block_bytes)
(quantizer.cpp:19); f8e5m2 weights carry an F32 `.scale` sibling of shape ceil(rows/128) x ceil(cols/128).

This module is host tooling (checkpoint minting, tests, bench); the C++ loader the engine uses lives in
csrc/host/dseek_loader.cpp.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Tuple

import numpy as np

# dtype strings: src/codec.cpp:85-107
_DTYPES = {
    "F32": np.dtype("<f4"),
    "F16": np.dtype("<f2"),
    "BF16": np.dtype("<u2"),
    "F8_E5M2": np.dtype("u1"),
    "F8_E4M3": np.dtype("u1"),
    "I32": np.dtype("<i4"),
    "I16": np.dtype("<i2"),
    "I8": np.dtype("i1"),
    "U8": np.dtype("u1"),
}

Q2K_BLOCK_BYTES = 84   # src/quant.h:41-52
Q3K_BLOCK_BYTES = 110  # src/quant.h:70-76
QK_K = 256


class DseekTensor:
    __slots__ = ("name", "dtype", "shape", "data")

    def __init__(self, name: str, dtype: str, shape: Tuple[int, ...], data: np.ndarray):
        self.name, self.dtype, self.shape, self.data = name, dtype, tuple(shape), data

    def __repr__(self):
        return f"DseekTensor({self.name}, {self.dtype}, {self.shape})"


def write_shard(path: str, tensors: Dict[str, Tuple[str, np.ndarray]], metadata: Dict[str, str] | None = None):
    """tensors: name -> (dtype string, ndarray whose raw bytes are the payload, shape = ndarray.shape)."""
    header = {}
    if metadata is not None:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    off = 0
    order = sorted(tensors.keys())
    for name in order:
        dt, arr = tensors[name]
        nbytes = arr.nbytes
        assert nbytes == int(np.prod(arr.shape, dtype=np.int64)) * _DTYPES[dt].itemsize, name
        header[name] = {"dtype": dt, "shape": list(arr.shape), "data_offsets": [off, off + nbytes]}
        off += nbytes
    hjson = json.dumps(header, separators=(",", ":")).encode("utf-8")
    pad = (-len(hjson)) % 8
    hjson += b" " * pad
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hjson)))
        f.write(hjson)
        for name in order:
            arr = np.ascontiguousarray(tensors[name][1])
            f.write(arr.tobytes() if arr.nbytes < (1 << 26) else memoryview(arr).cast("B"))


def read_dir(dirname: str):
    """Returns (metadata dict, {name: DseekTensor}) with zero-copy memmaps."""
    files = sorted(os.path.join(dirname, f) for f in os.listdir(dirname))
    if not files:
        raise FileNotFoundError(f"no shards in {dirname}")
    metadata = None
    tensors: Dict[str, DseekTensor] = {}
    for idx, path in enumerate(files):
        mm = np.memmap(path, dtype=np.uint8, mode="r")
        (hlen,) = struct.unpack("<Q", mm[:8].tobytes())
        header = json.loads(mm[8 : 8 + hlen].tobytes().decode("utf-8"))
        base = 8 + hlen
        for name, val in header.items():
            if name == "__metadata__":
                if idx == 0:
                    metadata = val
                continue
            b, e = val["data_offsets"]
            dt = _DTYPES[val["dtype"]]
            arr = mm[base + b : base + e].view(dt).reshape(val["shape"])
            tensors[name] = DseekTensor(name, val["dtype"], tuple(val["shape"]), arr)
    if metadata is None:
        raise ValueError("first shard has no __metadata__")
    return metadata, tensors
