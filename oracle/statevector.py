"""Oracle: exact statevector simulation (<= ~22 qubits).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pin (b) of SURVEY.md
section 8(c): the *unrouted* circuit (XXPhase applied directly on (a, b))
is simulated exactly, which validates routing + MPS + truncation jointly.
Qubit i is tensor axis i (qubit 0 most significant), matching MPS site i
(KernelPkg.jl:50-60 maps qubit q to site 1+q).
"""

from __future__ import annotations

import numpy as np

from .ansatz import bind_gate_list, gate_matrix, unrouted_gate_list


def apply_gate_sv(psi: np.ndarray, mat: np.ndarray, qubits) -> np.ndarray:
    n = psi.ndim
    if len(qubits) == 1:
        (q,) = qubits
        psi = np.tensordot(mat, psi, axes=([1], [q]))
        return np.moveaxis(psi, 0, q)
    q0, q1 = qubits
    m4 = mat.reshape(2, 2, 2, 2)
    psi = np.tensordot(m4, psi, axes=([2, 3], [q0, q1]))
    return np.moveaxis(psi, [0, 1], [q0, q1])


def run_gates_sv(n: int, bound_gates) -> np.ndarray:
    psi = np.zeros((2,) * n, dtype=np.complex128)
    psi[(0,) * n] = 1.0
    for name, qubits, params in bound_gates:
        mat = gate_matrix(name, params[0] if params else None)
        psi = apply_gate_sv(psi, mat, qubits)
    return psi.reshape(-1)


def statevector_for_data(num_qubits, reps, gamma, entanglement_map, x, hadamard_init=True):
    gates = unrouted_gate_list(num_qubits, reps, gamma, entanglement_map, hadamard_init)
    return run_gates_sv(num_qubits, bind_gate_list(gates, x))


def statevector_gram(num_qubits, reps, gamma, entanglement_map, X, Y=None, hadamard_init=True):
    """K[y, x] = |<psi(Y[y])|psi(X[x])>|^2 from exact statevectors."""
    sx = np.stack([statevector_for_data(num_qubits, reps, gamma, entanglement_map, x, hadamard_init) for x in X])
    sy = sx if Y is None else np.stack(
        [statevector_for_data(num_qubits, reps, gamma, entanglement_map, y, hadamard_init) for y in Y])
    ov = sy.conj() @ sx.T
    return (ov * ov.conj()).real
