"""Writes the golden fixtures in this directory.

The reference ships no golden vectors and cannot be executed offline (SURVEY.md sections 4, 8(c)),
so the fixtures are produced by the oracle (CPU restatement, ITensors semantics, cutoff 1e-16) and,
where the qubit count allows, by the exact statevector of the unrouted circuit.  Inputs are the
synthetic Elliptic-shaped features of oracle.synth (seeded).

    python tests/golden/make_golden.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from oracle.gram_ref import gram_from_mps, simulate_batch  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent
CASES = [
    # name, n, r, gamma, d, Nx, Ny, exact?
    ("c1_10q_r2_g0.5_d1", 10, 2, 0.5, 1, 40, 8, True),     # BASELINE config 1 (40x40 train + 8x40 test)
    ("c2_20q_r2_g0.5_d1", 20, 2, 0.5, 1, 12, 0, True),     # BASELINE config 2 shape, statevector cross-check
    ("c3_50q_r2_g0.1_d2", 50, 2, 0.1, 2, 10, 4, False),    # BASELINE config 3 shape, gamma 0.1
    ("c3_50q_r2_g1.0_d2", 50, 2, 1.0, 2, 8, 0, False),     # BASELINE config 3 shape, gamma 1.0 (max chi)
    ("d3_12q_r2_g0.7_d3", 12, 2, 0.7, 3, 6, 3, True),      # chi up to 32
]

for name, n, r, g, d, nx, ny, exact in CASES:
    emap = oracle.entanglement_graph(n, d)
    X = oracle.synthetic_features(nx, n, 0)
    xm = simulate_batch(n, r, g, emap, X)
    out = dict(n=n, r=r, gamma=g, d=d, X=X, chi_X=np.array([[1] + m.bond_dims() + [1] for m in xm]))
    if ny:
        Y = oracle.synthetic_features(ny, n, 1)
        ym = simulate_batch(n, r, g, emap, Y)
        out["Y"] = Y
        out["K_oracle"] = gram_from_mps(xm, ym)
        if exact:
            out["K_exact"] = oracle.statevector_gram(n, r, g, emap, X, Y)
    else:
        out["K_oracle"] = gram_from_mps(xm)
        if exact:
            out["K_exact"] = oracle.statevector_gram(n, r, g, emap, X)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, out["K_oracle"].shape, "max chi", out["chi_X"].max())
