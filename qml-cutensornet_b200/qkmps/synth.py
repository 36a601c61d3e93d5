"""Synthetic Elliptic-shaped inputs (the dataset is not available offline):
features scaled to [0, 2] per column, the range main.py:138-140's MinMaxScaler((0, 2)) produces."""

import numpy as np


def synthetic_features(n_points: int, n_features: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_points, n_features))
    lo, hi = x.min(axis=0), x.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    return np.ascontiguousarray(2.0 * (x - lo) / span)


def entanglement_graph(nq: int, nn: int):
    """Linear distance-<=nn entanglement map in the order of the reference's main.py:21-45
    (per distance: greedy non-overlapping pairs, then the pairs starting on their right qubits)."""
    pairs = []
    for d in range(1, nn + 1):
        rights = set()
        for i in range(nq):
            if i not in rights and i + d < nq:
                pairs.append((i, i + d))
                rights.add(i + d)
        pairs.extend((i, i + d) for i in rights if i + d < nq)
    return pairs
