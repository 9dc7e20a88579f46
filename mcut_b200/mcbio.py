"""MCB1 named-array container (see oracle/mcb_io.hpp): the exchange format between the python
test/bench drivers and the C++ harnesses.  Plain numpy, no torch."""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_DTYPES = [np.uint8, np.uint32, np.int32, np.uint64, np.float32, np.float64, np.int64]
_CODES = {np.dtype(d): i for i, d in enumerate(_DTYPES)}


def write_mcb(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as fp:
        fp.write(b"MCB1")
        fp.write(struct.pack("<I", len(arrays)))
        for name in sorted(arrays):
            a = np.ascontiguousarray(arrays[name])
            if a.dtype not in _CODES:
                raise TypeError(f"mcb: unsupported dtype {a.dtype} for {name}")
            nb = name.encode()
            fp.write(struct.pack("<I", len(nb)))
            fp.write(nb)
            fp.write(struct.pack("<II", _CODES[a.dtype], a.ndim))
            if a.ndim:
                fp.write(struct.pack(f"<{a.ndim}Q", *a.shape))
            fp.write(a.tobytes())


def read_mcb(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as fp:
        buf = fp.read()
    if buf[:4] != b"MCB1":
        raise ValueError(f"mcb: bad magic in {path}")
    (n,) = struct.unpack_from("<I", buf, 4)
    off = 8
    for _ in range(n):
        (nl,) = struct.unpack_from("<I", buf, off)
        off += 4
        name = buf[off:off + nl].decode()
        off += nl
        code, nd = struct.unpack_from("<II", buf, off)
        off += 8
        dims = struct.unpack_from(f"<{nd}Q", buf, off) if nd else ()
        off += 8 * nd
        dt = np.dtype(_DTYPES[code])
        cnt = int(np.prod(dims)) if nd else 1
        out[name] = np.frombuffer(buf, dtype=dt, count=cnt, offset=off).reshape(dims).copy()
        off += cnt * dt.itemsize
    return out
