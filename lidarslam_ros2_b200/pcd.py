"""Minimal reader for binary PCD files (the two fixtures of Thirdparty/ndt_omp_ros2/data/, FIELDS x y z intensity,
DATA binary) — what apps/align.cpp:55-62 loads with pcl::io::loadPCDFile."""
from __future__ import annotations

import numpy as np


def load_pcd(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    marker = b"DATA binary\n"
    k = data.index(marker) + len(marker)
    header = data[:k].decode("ascii", "replace").splitlines()
    fields = next(l for l in header if l.startswith("FIELDS")).split()[1:]
    sizes = [int(v) for v in next(l for l in header if l.startswith("SIZE")).split()[1:]]
    n = int(next(l for l in header if l.startswith("POINTS")).split()[1])
    if any(s != 4 for s in sizes):
        raise ValueError("only 4-byte fields supported")
    return np.frombuffer(data[k:k + n * 4 * len(fields)], dtype=np.float32).reshape(n, len(fields)).copy()
