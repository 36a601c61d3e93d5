"""Ansatz definition without pytket / sympy: same constructor, attributes and error behaviour as the
reference's ``KernelStateAnsatz`` (gpu_backend/kernel_state_ansatz.py:16-103,
cpu_backend/kernel_state_ansatz.py:20-131), backed by a plain symbolic gate list.

A symbolic gate is ``(name, qubits, param)`` with ``param`` in
    None | ("lin", i, coeff) | ("prod", a, b, coeff) | ("const", alpha)
meaning alpha = coeff*f_i, coeff*(1-f_a)*(1-f_b) or a constant; alpha is in half-turns (TKET).
"""

from __future__ import annotations

import math


class FeatureSymbol(str):
    """Stands in for ``sympy.Symbol('f_i')``; only identity and ``.name`` are ever used."""

    @property
    def name(self):
        return str(self)


class SymbolicCircuit:
    """What the reference keeps in ``ansatz.ansatz_circ`` (a pytket Circuit): here a gate list."""

    def __init__(self, n_qubits: int, gates=None):
        self.n_qubits = int(n_qubits)
        self.gates = list(gates) if gates is not None else []

    def add(self, name, qubits, param=None):
        self.gates.append((name, tuple(int(q) for q in qubits), param))

    def get_commands(self):
        return list(self.gates)

    def copy(self):
        return SymbolicCircuit(self.n_qubits, self.gates)

    @property
    def n_gates(self):
        return len(self.gates)


def eval_param(param, values):
    if param is None:
        return None
    kind = param[0]
    if kind == "lin":
        return param[2] * values[param[1]]
    if kind == "prod":
        return param[3] * (1 - values[param[1]]) * (1 - values[param[2]])
    if kind == "const":
        return param[1]
    raise RuntimeError(f"bad gate parameter {param!r}")


class BoundCircuit(SymbolicCircuit):
    """Circuit with every symbol substituted (what ``circuit_for_data`` returns on the GPU backend)."""

    def as_tuples(self):
        return [(name, list(q), [] if p is None else [float(p[1])]) for name, q, p in self.gates]


class KernelStateAnsatzBase:
    """Symbolic ansatz circuit U(x)|0>; see the module docstring for the reference lines mirrored."""

    def __init__(self, num_qubits, reps, gamma, entanglement_map, hadamard_init=True):
        self.one_q_symbol_list = []
        self.two_q_symbol_list = []
        self.num_qubits = int(num_qubits)
        self.reps = int(reps)
        self.gamma = float(gamma)
        self.entanglement_map = [(int(a), int(b)) for (a, b) in entanglement_map]
        self.hadamard_init = bool(hadamard_init)
        self.feature_symbol_list = [FeatureSymbol("f_" + str(i)) for i in range(self.num_qubits)]

        circ = SymbolicCircuit(self.num_qubits)
        if self.hadamard_init:
            for q in range(self.num_qubits):
                circ.add("H", (q,))
        for _ in range(self.reps):
            rz_coeff = (2 / math.pi) * self.gamma
            for q in range(self.num_qubits):
                circ.add("Rz", (q,), ("lin", q, rz_coeff))
            for (a, b) in self.entanglement_map:
                lo, hi = (a, b) if a < b else (b, a)
                # eager routing: walk qubit `lo` next to `hi`, interact, walk it back
                for q in range(lo, hi - 1):
                    circ.add("SWAP", (q, q + 1))
                circ.add("XXPhase", (hi - 1, hi), ("prod", a, b, self.gamma * self.gamma))
                for q in range(hi - 2, lo - 1, -1):
                    circ.add("SWAP", (q, q + 1))
        self.ansatz_circ = circ

    def _bind(self, feature_values) -> BoundCircuit:
        if len(feature_values) != len(self.feature_symbol_list):
            raise RuntimeError("The number of values must match the number of symbols.")
        vals = [float(v) for v in feature_values]
        bound = []
        for name, qubits, param in self.ansatz_circ.gates:
            a = eval_param(param, vals)
            bound.append((name, qubits, None if a is None else ("const", a)))
        return BoundCircuit(self.num_qubits, bound)


def _max_matching(edges, left_nodes):
    """Bipartite maximum matching (Kuhn); edges: dict left -> list of right."""
    match_r = {}

    def try_aug(u, seen):
        for v in edges.get(u, ()):
            if v in seen:
                continue
            seen.add(v)
            if v not in match_r or try_aug(match_r[v], seen):
                match_r[v] = u
                return True
        return False

    return sum(1 for u in left_nodes if try_aug(u, set()))


def structural_chi_bound(num_qubits, reps, entanglement_map) -> int:
    """Upper bound on the final Schmidt rank over every cut: 2^(reps * minimum vertex cover of the
    XX gates crossing the cut), capped by the chain-edge bound."""
    best = 1
    for cut in range(num_qubits - 1):           # between sites cut and cut+1
        edges, left = {}, set()
        for (a, b) in entanglement_map:
            lo, hi = (a, b) if a < b else (b, a)
            if lo <= cut < hi:
                edges.setdefault(lo, []).append(hi)
                left.add(lo)
        cover = _max_matching(edges, sorted(left))   # Koenig: min vertex cover = max matching
        edge = min(cut + 1, num_qubits - cut - 1)
        best = max(best, 2 ** min(reps * cover, edge, 30))
    return best


def expected_chi(gamma: float, structural: int, truncation_error: float = 1e-16, n_terms: int = 1) -> int:
    """Cheap estimate of the bond dimension the truncation rule will keep, used to pick the first bond cap.

    An XXPhase interaction has angle theta <= (pi/2) gamma^2 (features lie in [0, 2]); every additional
    Schmidt vector it creates carries a squared weight ~ theta^2 relative to the previous one, so weights
    fall below ``truncation_error`` after about log(truncation_error) / log(theta^2) vectors.  ``n_terms`` =
    repetitions x distance counts how many interactions pile up on one cut; the estimate is only trusted for
    shallow circuits (n_terms <= 4), where it was checked against measured bond dimensions -- a wrong guess
    costs a partial run plus a re-run of the datapoints that hit the cap (engine._simulate_shard).
    """
    theta = (math.pi / 2) * gamma * gamma
    if theta <= 0.0:
        return 1          # gamma = 0: no interaction at all, product state
    if theta >= 1.0 or truncation_error <= 0 or n_terms > 4:
        return structural
    k = math.log(max(truncation_error, 1e-300)) / math.log(theta * theta)
    return int(min(structural, 1 + math.ceil(k)))
