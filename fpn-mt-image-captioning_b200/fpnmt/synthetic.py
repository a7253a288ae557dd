"""Deterministic synthetic inputs (numpy only): images in the value range of the reference's preprocessing
(/root/reference/dataset.py:19-26, `x/127.5 - 1` -> [-1, 1], NHWC float32).

`uniform_images` is the throughput workload of BASELINE.md (i.i.d. U(-1,1) pixels).  `structured_images` are smooth
random colour fields plus a little noise: unlike i.i.d. pixels they differ from one another after the backbone's
spatial averaging, so parity tests on them exercise the input-dependent part of every stage.
"""
from __future__ import annotations

import numpy as np


def uniform_images(n: int, size: int = 512, seed: int = 1234) -> np.ndarray:
    r = np.random.default_rng(seed)
    return (r.random((n, size, size, 3), dtype=np.float32) * 2 - 1).astype(np.float32)


def structured_images(n: int, size: int = 512, seed: int = 1, grid: int = 6, noise: float = 0.15) -> np.ndarray:
    r = np.random.default_rng(seed)
    lo = r.uniform(-1, 1, (n, grid, grid, 3)).astype(np.float32)
    # separable linear interpolation of the coarse grid to size x size (align-corners style), pure numpy
    pos = np.linspace(0, grid - 1, size, dtype=np.float32)
    i0 = np.minimum(np.floor(pos).astype(np.int64), grid - 2)
    f = (pos - i0).astype(np.float32)
    rows = lo[:, i0] * (1 - f)[None, :, None, None] + lo[:, i0 + 1] * f[None, :, None, None]
    img = rows[:, :, i0] * (1 - f)[None, None, :, None] + rows[:, :, i0 + 1] * f[None, None, :, None]
    img = img * (1 - noise) + r.uniform(-noise, noise, (n, size, size, 3)).astype(np.float32)
    return np.clip(img, -1, 1).astype(np.float32)
