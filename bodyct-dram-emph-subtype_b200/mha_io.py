"""Native MetaImage (.mha) reader/writer: the subset SimpleITK is used for in the reference
(dataset.py:49-55 sitk.ReadImage/GetArrayFromImage, utils.py:87-104 ImageFileWriter with compression).

A .mha file is an ASCII `Key = Value` header terminated by `ElementDataFile = LOCAL`, followed by the
raw (optionally zlib-compressed) voxels with x fastest.  Arrays are returned/accepted as numpy
[z, y, x]; spacing/origin are in ITK order (x, y, z) and `direction` is the row-major 3x3 matrix sitk's
GetDirection() returns.  In the file, MetaIO stores `TransformMatrix` row i = direction cosines of image axis i —
the TRANSPOSE of that matrix (itkMetaImageIO.cxx) — so it is transposed on read and on write; header key order,
`CompressedDataSize` and `AnatomicalOrientation` follow what ITK's writer emits (pinned by the hand-built ITK-layout
fixtures tests/golden/itk_style_*.mha, oracle/make_mha_fixture.py).
"""
import zlib
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_MET_TO_NP = {
    "MET_CHAR": np.int8, "MET_UCHAR": np.uint8, "MET_SHORT": np.int16, "MET_USHORT": np.uint16,
    "MET_INT": np.int32, "MET_UINT": np.uint32, "MET_LONG": np.int32, "MET_ULONG": np.uint32,
    "MET_LONG_LONG": np.int64, "MET_ULONG_LONG": np.uint64, "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64,
}
_NP_TO_MET = {np.dtype(v): k for k, v in _MET_TO_NP.items() if "LONG" not in k or "LONG_LONG" in k}


def read_mha(path):
    """Returns (array [z,y,x], meta dict with spacing, origin, direction as tuples in ITK order)."""
    with open(path, "rb") as f:
        blob = f.read()
    header = {}
    pos = 0
    while True:
        end = blob.index(b"\n", pos)
        line = blob[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line:
            continue
        key, _, val = line.partition("=")
        header[key.strip()] = val.strip()
        if key.strip() == "ElementDataFile":
            break
    if header.get("ElementDataFile") != "LOCAL":
        raise ValueError(f"{path}: only single-file MetaImage (ElementDataFile = LOCAL) is supported")
    if header.get("ObjectType", "Image") != "Image":
        raise ValueError(f"{path}: not an image")
    ndims = int(header.get("NDims", 3))
    dims = [int(v) for v in header["DimSize"].split()]
    if ndims != 3 or len(dims) != 3:
        raise ValueError(f"{path}: expected a 3-D image, got NDims={ndims}")
    if int(header.get("ElementNumberOfChannels", 1)) != 1:
        raise ValueError(f"{path}: multi-channel images are not supported")
    dtype = np.dtype(_MET_TO_NP[header["ElementType"]])
    msb = header.get("BinaryDataByteOrderMSB", header.get("ElementByteOrderMSB", "False")).lower() == "true"
    data = blob[pos:]
    if header.get("CompressedData", "False").lower() == "true":
        data = zlib.decompress(data)
    count = dims[0] * dims[1] * dims[2]
    arr = np.frombuffer(data, dtype=dtype.newbyteorder(">" if msb else "<"), count=count)
    arr = arr.astype(dtype, copy=True).reshape(dims[2], dims[1], dims[0])
    floats = lambda key, default: tuple(float(v) for v in header.get(key, default).split())  # noqa: E731
    t = floats("TransformMatrix", header.get("Rotation", header.get("Orientation", "1 0 0 0 1 0 0 0 1")))
    meta = {
        "spacing": floats("ElementSpacing", "1 1 1"),
        "origin": floats("Offset", header.get("Position", header.get("Origin", "0 0 0"))),
        "direction": _transpose3(t),
    }
    return arr, meta


def _transpose3(m):
    m = tuple(m)
    if len(m) != 9:
        raise ValueError(f"expected a 3x3 direction matrix, got {len(m)} values")
    return (m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8])


def write_mha(path, arr, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0),
              direction=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0), compress=True):
    arr = np.ascontiguousarray(arr)
    if arr.ndim != 3:
        raise ValueError("write_mha: expected a 3-D [z,y,x] array")
    if arr.dtype not in _NP_TO_MET:
        raise TypeError(f"write_mha: unsupported dtype {arr.dtype}")
    raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
    payload = zlib.compress(raw, 2) if compress else raw
    fmt = lambda vals: " ".join(repr(float(v)) if float(v) != int(v) else str(int(v)) for v in vals)  # noqa: E731
    lines = [
        "ObjectType = Image", "NDims = 3", "BinaryData = True", "BinaryDataByteOrderMSB = False",
        f"CompressedData = {'True' if compress else 'False'}",
    ]
    if compress:
        lines.append(f"CompressedDataSize = {len(payload)}")
    lines += [
        f"TransformMatrix = {fmt(_transpose3(float(v) for v in direction))}", f"Offset = {fmt(origin)}",
        "CenterOfRotation = 0 0 0", "AnatomicalOrientation = RAI", f"ElementSpacing = {fmt(spacing)}",
        f"DimSize = {arr.shape[2]} {arr.shape[1]} {arr.shape[0]}", f"ElementType = {_NP_TO_MET[arr.dtype]}",
        "ElementDataFile = LOCAL",
    ]
    with open(path, "wb") as f:
        f.write(("\n".join(lines) + "\n").encode("ascii"))
        f.write(payload)


class BackgroundWriter:
    """`--workers N` of the processor, output side: file writes (zlib, which releases the GIL) run on N threads while
    the predict loop goes on.  At most 2 N calls are in flight (each holds a full-size volume); an exception of a
    write surfaces at the next `submit` or at `close()`.  N = 0 runs every call at once in the caller's thread."""

    def __init__(self, workers=0):
        self.pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="mha-write") if workers > 0 else None
        self.limit = 2 * max(1, workers)
        self.pending = deque()

    def submit(self, fn, *args, **kwargs):
        if self.pool is None:
            fn(*args, **kwargs)
            return
        while len(self.pending) >= self.limit or (self.pending and self.pending[0].done()):
            self.pending.popleft().result()
        self.pending.append(self.pool.submit(fn, *args, **kwargs))

    def close(self):
        try:
            while self.pending:
                self.pending.popleft().result()
        finally:
            if self.pool is not None:
                self.pool.shutdown(wait=True)
                self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
