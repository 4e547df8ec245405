""".fvecs / .ivecs readers and writers — same on-disk format as the reference's
ReadData / WriteData (include/util.h:52-80,149-168): per row [int32 d][d x 4 bytes]."""
from __future__ import annotations

import numpy as np


def read_vecs(path: str, dtype) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.int32)
    if raw.size == 0:
        return np.zeros((0, 0), dtype=dtype)
    d = int(raw[0])
    rows = raw.size // (d + 1)           # util.h:65  num = fsize / (dim + 1) / 4
    return np.ascontiguousarray(raw[: rows * (d + 1)].reshape(rows, d + 1)[:, 1:]).view(dtype)


def read_fvecs(path: str) -> np.ndarray:
    return read_vecs(path, np.float32)


def read_ivecs(path: str) -> np.ndarray:
    return read_vecs(path, np.uint32)


def write_vecs(path: str, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a)
    assert a.dtype.itemsize == 4 and a.ndim == 2
    n, d = a.shape
    out = np.empty((n, d + 1), dtype=np.int32)
    out[:, 0] = d
    out[:, 1:] = a.view(np.int32)
    out.tofile(path)
