"""Oracle: entanglement map, ansatz circuit with SWAP routing, gate matrices.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  No pytket / sympy: the
circuit is a plain list of symbolic gates

    (name, qubits, param)

with ``param`` one of
    None                          H, SWAP
    ("lin",  i, coeff)            alpha = coeff * x[i]                 (Rz, Rx)
    ("prod", a, b, coeff)         alpha = coeff * (1-x[a]) * (1-x[b])  (XXPhase, ZZPhase)
    ("const", alpha)              fixed angle
``alpha`` is in half-turns, TKET convention (reference
KernelPkg/src/KernelPkg.jl:8-42: theta = pi*alpha/2).
"""

from __future__ import annotations

import math

import numpy as np


def entanglement_graph(nq: int, nn: int) -> list[tuple[int, int]]:
    """Ordered qubit pairs for the linear, distance<=nn entanglement map.

    Restates reference main.py:21-45.  For every distance d the gates come in
    two sub-layers: first the greedy non-overlapping pairs (i, i+d), then the
    pairs whose left qubit was a right qubit of the first sub-layer.  The
    reference iterates a Python ``set`` for the second sub-layer
    (main.py:41); we do the same so that the order is the reference's order
    on the same interpreter (ascending for every BASELINE config).
    """
    pairs: list[tuple[int, int]] = []
    for d in range(1, nn + 1):
        busy: set[int] = set()
        for i in range(nq):
            if i not in busy and i + d < nq:
                pairs.append((i, i + d))
                busy.add(i + d)
        for i in busy:
            if i + d < nq:
                pairs.append((i, i + d))
    return pairs


def ansatz_gate_list(num_qubits, reps, gamma, entanglement_map, hadamard_init=True):
    """Symbolic routed circuit of KernelStateAnsatz.__init__.

    Restates reference gpu_backend/kernel_state_ansatz.py:53-90 (identical to
    cpu_backend/kernel_state_ansatz.py:57-94):

      * H on every qubit (if hadamard_init)                       gpu:53-55
      * reps x [ Rz((2/pi)*gamma*f_i) on every qubit,             gpu:57-60
                 XXPhase(gamma^2 (1-f_a)(1-f_b)) per map pair ]   gpu:62-66
      * eager routing: SWAP(q,q+1) for q=q0..q1-2, XXPhase on
        (q1-1,q1), SWAPs in reverse                               gpu:78-88

    Commands are emitted in insertion order, which is one valid topological
    order of the circuit (pytket's get_commands() may pick another; every XX
    gate of a layer commutes with every other).
    """
    gates = []
    if hadamard_init:
        for i in range(num_qubits):
            gates.append(("H", (i,), None))
    for _ in range(reps):
        for i in range(num_qubits):
            gates.append(("Rz", (i,), ("lin", i, (2.0 / np.pi) * gamma)))
        for (a, b) in entanglement_map:
            q0, q1 = (min(a, b), max(a, b))
            param = ("prod", a, b, gamma * gamma)
            for q in range(q0, q1 - 1):
                gates.append(("SWAP", (q, q + 1), None))
            gates.append(("XXPhase", (q1 - 1, q1), param))
            for q in reversed(range(q0, q1 - 1)):
                gates.append(("SWAP", (q, q + 1), None))
    return gates


def unrouted_gate_list(num_qubits, reps, gamma, entanglement_map, hadamard_init=True):
    """Same circuit with XXPhase acting directly on (a, b) -- for the statevector check."""
    gates = []
    if hadamard_init:
        for i in range(num_qubits):
            gates.append(("H", (i,), None))
    for _ in range(reps):
        for i in range(num_qubits):
            gates.append(("Rz", (i,), ("lin", i, (2.0 / np.pi) * gamma)))
        for (a, b) in entanglement_map:
            gates.append(("XXPhase", (a, b), ("prod", a, b, gamma * gamma)))
    return gates


def eval_param(param, x) -> float | None:
    if param is None:
        return None
    kind = param[0]
    if kind == "lin":
        return param[2] * x[param[1]]
    if kind == "prod":
        return param[3] * (1.0 - x[param[1]]) * (1.0 - x[param[2]])
    if kind == "const":
        return param[1]
    raise RuntimeError(f"bad param {param!r}")


def bind_gate_list(gates, x):
    """circuit_for_data of the CPU backend (cpu:96-131): list of (name, [q..], [alpha])."""
    x = np.asarray(x, dtype=np.float64)
    out = []
    for name, qubits, param in gates:
        a = eval_param(param, x)
        out.append((name, list(qubits), [] if a is None else [float(a)]))
    return out


_SQ = 1.0 / math.sqrt(2.0)


def gate_matrix(name: str, alpha: float | None = None) -> np.ndarray:
    """Gate unitaries, TKET convention as written in KernelPkg/src/KernelPkg.jl:8-42.

    Two-qubit matrices are indexed [(L,R),(l,r)] with the first listed qubit as
    the more significant bit (ITensors ``op`` on (s1, s2), KernelPkg.jl:56-60).
    """
    if name == "H":
        return np.array([[_SQ, _SQ], [_SQ, -_SQ]], dtype=np.complex128)
    if name == "SWAP":
        m = np.zeros((4, 4), dtype=np.complex128)
        m[0, 0] = m[1, 2] = m[2, 1] = m[3, 3] = 1.0
        return m
    th = math.pi * alpha / 2.0
    c, s = math.cos(th), math.sin(th)
    if name == "Rx":  # KernelPkg.jl:8-14
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "Rz":  # KernelPkg.jl:16-22
        return np.array([[complex(c, -s), 0], [0, complex(c, s)]], dtype=np.complex128)
    if name == "XXPhase":  # KernelPkg.jl:24-32
        m = np.zeros((4, 4), dtype=np.complex128)
        for i in range(4):
            m[i, i] = c
            m[i, 3 - i] = -1j * s
        return m
    if name == "ZZPhase":  # KernelPkg.jl:34-42
        e_m, e_p = complex(c, -s), complex(c, s)
        return np.diag([e_m, e_p, e_p, e_m]).astype(np.complex128)
    raise RuntimeError("KernelPkg error: Unrecognised gate.")  # KernelPkg.jl:62
