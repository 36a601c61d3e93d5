"""Oracle: Gram tiles and the full kernel matrix on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``compute_tile`` restates KernelPkg/src/KernelPkg.jl:75-112 (simulate every x
and every y circuit of the tile, then all pairs ``abs(inner(y, x))^2``);
``gram_matrix`` restates the tiling / symmetry logic of
cpu_backend/kernel_state_ansatz.py:176-203,243-274 without MPI, with the
(sensible) difference that every circuit is simulated once instead of once
per tile -- results are identical because the simulation is deterministic.
"""

from __future__ import annotations

import time

import numpy as np

from .ansatz import ansatz_gate_list, bind_gate_list
from .mps_ref import mps_inner, simulate_mps


def compute_tile(n_qubits, x_circs, y_circs, cutoff, mode="itensors"):
    """Returns (tile[len(y), len(x)], x_chi, y_chi, x_time, y_time, vdot_time)."""
    tile = np.zeros((len(y_circs), len(x_circs)))
    x_mps, x_chi, x_time = [], [], []
    for circ in x_circs:
        t0 = time.perf_counter()
        m = simulate_mps(n_qubits, circ, cutoff, mode)
        x_time.append(time.perf_counter() - t0)
        x_chi.append(m.max_chi())
        x_mps.append(m)
    y_mps, y_chi, y_time = [], [], []
    for circ in y_circs:
        t0 = time.perf_counter()
        m = simulate_mps(n_qubits, circ, cutoff, mode)
        y_time.append(time.perf_counter() - t0)
        y_chi.append(m.max_chi())
        y_mps.append(m)
    vdot_time = []
    for i, ym in enumerate(y_mps):
        for j, xm in enumerate(x_mps):
            t0 = time.perf_counter()
            tile[i, j] = abs(mps_inner(ym, xm)) ** 2
            vdot_time.append(time.perf_counter() - t0)
    return tile, x_chi, y_chi, x_time, y_time, vdot_time


def simulate_batch(num_qubits, reps, gamma, entanglement_map, X, cutoff=1e-16, mode="itensors",
                   hadamard_init=True, chi=None):
    gates = ansatz_gate_list(num_qubits, reps, gamma, entanglement_map, hadamard_init)
    return [simulate_mps(num_qubits, bind_gate_list(gates, x), cutoff, mode, chi=chi) for x in X]


def gram_from_mps(x_mps, y_mps=None):
    """K[y, x] = |<y|x>|^2; symmetric fill when y_mps is None (cpu:271-274)."""
    if y_mps is None:
        n = len(x_mps)
        k = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                v = abs(mps_inner(x_mps[j], x_mps[i])) ** 2
                k[j, i] = v
                k[i, j] = v
        return k
    k = np.zeros((len(y_mps), len(x_mps)))
    for i, ym in enumerate(y_mps):
        for j, xm in enumerate(x_mps):
            k[i, j] = abs(mps_inner(ym, xm)) ** 2
    return k


def gram_matrix(num_qubits, reps, gamma, entanglement_map, X, Y=None, cutoff=1e-16,
                mode="itensors", hadamard_init=True):
    """Full kernel matrix of shape [len(Y) or len(X), len(X)] (cpu:134-328 semantics)."""
    x_mps = simulate_batch(num_qubits, reps, gamma, entanglement_map, X, cutoff, mode, hadamard_init)
    y_mps = None if Y is None else simulate_batch(num_qubits, reps, gamma, entanglement_map, Y, cutoff, mode,
                                                  hadamard_init)
    return gram_from_mps(x_mps, y_mps)


def product_state_gram(reps, gamma, X, Y=None):
    """Closed form for an empty entanglement map (SURVEY.md 8(c) pin 1).

    psi(x) = (x)_k Rz-layers |+>  =>  K = prod_k cos^2(reps * gamma * (x_k - y_k)).
    """
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    d = Y[:, None, :] - X[None, :, :]
    return np.prod(np.cos(reps * gamma * d) ** 2, axis=2)
