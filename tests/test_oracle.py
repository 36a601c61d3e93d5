"""CPU tests pinning the oracle: literal gate matrices, closed form, exact
statevector, truncation rules on hand-made spectra, invariants, published-chi sanity, golden fixtures."""
import pathlib

import numpy as np
import pytest

import oracle
from oracle.ansatz import ansatz_gate_list, bind_gate_list, gate_matrix, unrouted_gate_list
from oracle.gram_ref import gram_from_mps, gram_matrix, product_state_gram, simulate_batch
from oracle.mps_ref import simulate_mps, truncate_itensors, truncate_pytket

GOLDEN = pathlib.Path(__file__).parent / "golden"
X2 = np.array([[0, 1], [1, 0]], dtype=complex)


def test_gate_matrices_literal():
    # KernelPkg.jl:8-42 known answers
    assert np.allclose(gate_matrix("XXPhase", 1.0), -1j * np.kron(X2, X2))
    assert np.allclose(gate_matrix("Rz", 1.0), np.diag([-1j, 1j]))
    assert np.allclose(gate_matrix("Rx", 1.0), -1j * X2)
    assert np.allclose(gate_matrix("ZZPhase", 1.0), np.diag([-1j, 1j, 1j, -1j]))
    assert np.allclose(gate_matrix("XXPhase", 0.0), np.eye(4))
    sw = gate_matrix("SWAP")
    v = np.arange(4.0)
    assert np.allclose(sw @ v, [0, 2, 1, 3])
    th = np.pi * 0.37 / 2
    assert np.allclose(gate_matrix("XXPhase", 0.37),
                       np.cos(th) * np.eye(4) - 1j * np.sin(th) * np.kron(X2, X2))
    with pytest.raises(RuntimeError):
        gate_matrix("CX", 0.1)


@pytest.mark.parametrize("n,d,n_pairs", [(10, 1, 9), (20, 1, 19), (50, 2, 97), (165, 4, 650), (100, 2, 197)])
def test_entanglement_graph_sizes(n, d, n_pairs):
    em = oracle.entanglement_graph(n, d)
    assert len(em) == n_pairs
    assert len(set(em)) == n_pairs
    assert all(0 <= a < b < n and b - a <= d for a, b in em)
    assert set(em) == {(i, i + k) for k in range(1, d + 1) for i in range(n - k)}


@pytest.mark.parametrize("n,r,d,n2q", [(10, 2, 1, 18), (20, 2, 1, 38), (50, 2, 2, 386), (100, 2, 2, 786), (165, 4, 4, 10360)])
def test_two_qubit_op_counts(n, r, d, n2q):
    gates = ansatz_gate_list(n, r, 0.5, oracle.entanglement_graph(n, d))
    assert sum(1 for g in gates if len(g[1]) == 2) == n2q
    assert all(q[1] == q[0] + 1 for _, q, _ in gates if len(q) == 2)   # routed: adjacent only


def test_closed_form_empty_map():
    X = oracle.synthetic_features(7, 9, 4)
    Y = oracle.synthetic_features(5, 9, 5)
    for r, g in [(1, 0.3), (2, 0.7), (3, 1.0)]:
        assert np.abs(gram_matrix(9, r, g, [], X) - product_state_gram(r, g, X)).max() < 1e-13
        assert np.abs(gram_matrix(9, r, g, [], X, Y) - product_state_gram(r, g, X, Y)).max() < 1e-13


@pytest.mark.parametrize("mode", ["itensors", "pytket"])
@pytest.mark.parametrize("n,r,g,d", [(8, 2, 0.5, 1), (10, 2, 1.0, 2), (11, 3, 0.4, 3)])
def test_mps_matches_statevector(n, r, g, d, mode):
    X = oracle.synthetic_features(5, n, 2)
    emap = oracle.entanglement_graph(n, d)
    K = gram_matrix(n, r, g, emap, X, mode=mode)
    Ksv = oracle.statevector_gram(n, r, g, emap, X)
    assert np.abs(K - Ksv).max() < 1e-8
    # state amplitudes too
    gates = ansatz_gate_list(n, r, g, emap)
    psi = simulate_mps(n, bind_gate_list(gates, X[0]), 1e-16, mode).to_statevector()
    sv = oracle.statevector_for_data(n, r, g, emap, X[0])
    assert np.abs(psi - sv).max() < 1e-7


def test_routing_equals_unrouted_unitary():
    n, r, g, d = 7, 2, 0.6, 3
    emap = oracle.entanglement_graph(n, d)
    x = oracle.synthetic_features(1, n, 0)[0]
    from oracle.statevector import run_gates_sv
    a = run_gates_sv(n, bind_gate_list(ansatz_gate_list(n, r, g, emap), x))
    b = run_gates_sv(n, bind_gate_list(unrouted_gate_list(n, r, g, emap), x))
    assert np.abs(a - b).max() < 1e-13


def test_invariants_and_commuting_reorder():
    n, r, g, d = 12, 2, 0.5, 2
    emap = oracle.entanglement_graph(n, d)
    X = oracle.synthetic_features(9, n, 1)
    K = gram_matrix(n, r, g, emap, X)
    assert np.array_equal(K, K.T)
    assert np.abs(np.diag(K) - 1).max() < 1e-12
    assert K.min() >= 0 and K.max() <= 1 + 1e-12
    assert np.linalg.eigvalsh(K).min() > -1e-10
    K2 = gram_matrix(n, r, g, list(reversed(emap)), X)      # every XX gate of a layer commutes
    assert np.abs(K - K2).max() < 1e-8
    bound = 2 ** (r * 2)                                        # 2^(r * cover(d)), cover(2) = 2
    assert max(m.max_chi() for m in simulate_batch(n, r, g, emap, X)) <= bound


def test_truncation_rules_on_handmade_spectra():
    # ITensors: walk from the tail while discarded + p <= cutoff * sum
    p = np.array([1.0, 1e-3, 4e-17, 3e-17, 2e-17])
    assert truncate_itensors(p, 1e-16)[0] == 2            # 2e-17+3e-17+4e-17 = 9e-17 <= 1e-16*(1.001)
    assert truncate_itensors(np.array([1.0, 6e-17, 5e-17]), 1e-16)[0] == 2   # 5e-17 ok, +6e-17 exceeds
    assert truncate_itensors(np.array([1.0, 0.0, 0.0]), 0.0)[0] == 1         # cutoff 0 drops exact zeros only
    assert truncate_itensors(np.array([1.0, 1e-40]), 0.0)[0] == 2
    assert truncate_itensors(np.array([0.5]), 1e-16)[0] == 1                 # single value: mindim 1
    assert truncate_itensors(np.array([1.0, 1.0, 1e-20, 1e-20]), 1e-16)[0] == 2   # ties in the tail
    assert truncate_itensors(np.array([1.0, 0.5, 0.25]), 1e-16, maxdim=2)[0] == 2
    assert truncate_itensors(np.array([1e-30, 1e-30]), 0.6)[0] == 1          # relative, not absolute
    # pytket: drop sigma < 1e-16, keep the shortest head reaching the fidelity
    k, kept = truncate_pytket(np.array([1.0, 1e-3, 1e-17]), 1 - 1e-16)
    assert k == 2 and kept == 1.0
    k, kept = truncate_pytket(np.array([0.8, 0.6]), 0.5)
    assert k == 1 and abs(kept - 0.64) < 1e-15
    k, kept = truncate_pytket(np.array([0.8, 0.6]), 0.64 + 1e-9)
    assert k == 2
    assert truncate_pytket(np.array([1.0, 0.9, 0.1]), 1.0, chi=2)[0] == 2


def test_published_chi_range():
    # runs/runtime_scaling/results.csv (n=165, r=2, d=1, gamma=0.1) reports avg max chi 2.0-2.03 on the real
    # Elliptic features; iid synthetic features give 3 under the 1e-16 rule (third squared Schmidt weight
    # ~ theta^4 ~ 1e-8 >> 1e-16).  Pin what is data-independent: 2 <= chi <= structural bound 2^(r*cover(1)) = 4.
    n = 165
    ms = simulate_batch(n, 2, 0.1, oracle.entanglement_graph(n, 1), oracle.synthetic_features(3, n, 0))
    assert all(2 <= m.max_chi() <= 4 for m in ms)


def test_golden_fixtures():
    """tests/golden/*.npz were written by tests/golden/make_golden.py (oracle + exact statevector)."""
    files = sorted(GOLDEN.glob("c*.npz")) + sorted((pathlib.Path(__file__).parent / "golden").glob("d*.npz"))
    assert files, "golden fixtures missing"
    for f in files:
        z = np.load(f)
        n, r, d = int(z["n"]), int(z["r"]), int(z["d"])
        g = float(z["gamma"])
        emap = oracle.entanglement_graph(n, d)
        K = gram_matrix(n, r, g, emap, z["X"], z["Y"] if "Y" in z.files else None)
        assert np.abs(K - z["K_oracle"]).max() < 1e-12, f.name
        if "K_exact" in z.files:
            assert np.abs(K - z["K_exact"]).max() < 1e-8, f.name
        chi = np.array([[1] + m.bond_dims() + [1] for m in simulate_batch(n, r, g, emap, z["X"])])
        assert np.array_equal(chi, z["chi_X"]), f.name


def test_pytket_rule_sums_in_one_order():
    """The pytket-rule restatement compares a running sum with the total accumulated in the same (largest-first) order,
    so it stops as soon as the remaining values no longer change the sum: rounding-noise values are not kept, a tail
    whose members each move the sum is."""
    from oracle.mps_ref import truncate_pytket
    s = np.array([1.0] + [1e-9] * 200)                    # squares 1e-18 each: below half an ulp of the sum
    k, kept = truncate_pytket(s, 1.0 - 1e-16)
    assert k == 1 and kept == 1.0
    s = np.array([1.0, 3e-8, 2e-8, 1e-9])                 # squares 9e-16 and 4e-16 move the sum, 1e-18 does not
    k, kept = truncate_pytket(s, 1.0 - 1e-16)
    assert k == 3
    k, kept = truncate_pytket(np.array([1.0, 0.5, 1e-17]), 1.0 - 1e-16)      # below value_of_zero: trimmed by the SVD itself
    assert k == 2
    k, kept = truncate_pytket(np.array([2.0, 1.0, 1.0, 0.5]), 1.0 - 0.2)       # looser fidelity: smallest prefix reaching 0.8
    assert k == 2 and abs(kept - 5.0 / 6.25) < 1e-15
    k, kept = truncate_pytket(np.array([2.0, 1.0, 1.0, 0.5]), 1.0 - 1e-16, chi=2)
    assert k == 2
