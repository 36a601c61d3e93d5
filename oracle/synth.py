"""Oracle-side synthetic inputs (the Elliptic dataset is not available offline).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  SURVEY.md 8(d): for
kernel-only runs ``X = minmax_[0,2](default_rng(seed).standard_normal((N, n)))``,
i.e. the value range main.py:138-140 (MinMaxScaler((0, 2)) fit on train)
produces.
"""

from __future__ import annotations

import numpy as np


def synthetic_features(n_points: int, n_features: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_points, n_features))
    lo, hi = x.min(axis=0), x.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    return np.ascontiguousarray(2.0 * (x - lo) / span)
