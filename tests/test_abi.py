"""CPU tests of the C ABI surface: libqkmps.so loads, exports every symbol include/qkmps.h declares,
the schedule compiler (host code) behaves, and compute entry points fail loudly without a GPU."""
import pathlib
import re

import numpy as np
import pytest

import oracle

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _declared_functions():
    text = (ROOT / "include" / "qkmps.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(qk):
    import ctypes
    L = ctypes.CDLL(str(qk.LIB_PATH))
    names = _declared_functions()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/qkmps.h but not exported"
    assert sorted(qk.EXPORTS) == names
    assert L.qk_version() == 100


def test_struct_layouts_match_header(qk):
    import ctypes
    assert ctypes.sizeof(qk.QkGate) == 32
    assert ctypes.sizeof(qk.QkOpView) == 32
    assert ctypes.sizeof(qk.QkPlanInfo) == 56


def _plan(qk, n, r, g, d, cap=16, mode=0, flags=0):
    gates = oracle.ansatz_gate_list(n, r, g, oracle.entanglement_graph(n, d))
    return qk.Plan(n, gates, mode, 1e-16, cap, flags), gates


def _replay(plan):
    """Replay a schedule: the orthogonality centre must sit on one of the two sites of every 2-site op."""
    centre = None
    two_q = []
    for kind, site, fa, fb, direction, coeff in plan.ops():
        if kind in (3, 4, 5):
            assert centre is None or centre in (site, site + 1), "centre not adjacent to the gate"
            centre = site if direction == 1 else site + 1
            two_q.append((kind, site, fa, fb))
        elif kind == 16:
            assert centre == site
            centre = site + 1
        elif kind == 17:
            assert centre == site
            centre = site - 1
        else:
            assert kind in (0, 1, 2)
    return two_q


@pytest.mark.parametrize("n,r,d,n2q", [(10, 2, 1, 18), (50, 2, 2, 386), (100, 2, 2, 786), (165, 4, 4, 10360)])
def test_plan_counts_and_schedule_invariants(qk, n, r, d, n2q):
    # literal order: the circuit's 2-qubit ops in the circuit's order, gauge moves in between
    plan, gates = _plan(qk, n, r, 0.5, d, flags=qk.QK_PLAN_LITERAL_ORDER)
    info = plan.info()
    assert info.n_qubits == n and info.n_gates == len(gates)
    assert info.n_ops_2q == n2q
    assert info.n_ops_1q == n * (r + 1)
    assert info.n_ops == info.n_ops_2q + info.n_ops_1q + info.n_moves
    assert info.n_moves > 0
    two_q = _replay(plan)
    ref = [(qk.GATE_KIND[nm], q[0]) for nm, q, _ in gates if len(q) == 2]
    assert [(k, s) for k, s, _, _ in two_q] == ref
    # reordering only: every run of (mutually commuting) XXPhase interactions applied as one sweep -> same
    # interactions, same routing per interaction, (almost) no gauge moves
    plan2 = qk.Plan(n, gates, 0, 1e-16, 16, qk.QK_PLAN_NO_FUSION)
    info2 = plan2.info()
    assert (info2.n_ops_2q, info2.n_ops_1q) == (n2q, n * (r + 1))
    assert info2.n_moves <= r
    two_q2 = _replay(plan2)
    assert sorted(two_q2) == sorted(two_q)
    xx = lambda ops: sorted((fa, fb) for k, s, fa, fb in ops if k == 3)   # noqa: E731
    assert xx(two_q2) == xx(two_q)
    # 1-qubit layers still separate the repetitions
    kinds = [k for k, *_ in plan2.ops()]
    assert kinds.count(1) == n * r and kinds.count(0) == n
    # default = reordering + fusion: gates that follow each other on one bond share an SVD, SWAP pairs cancel;
    # every interaction is still there exactly once, and fewer SVDs remain
    plan3, _ = _plan(qk, n, r, 0.5, d)
    info3 = plan3.info()
    ops3 = plan3.ops()
    assert xx([(k, s, fa, fb) for k, s, fa, fb, *_ in ops3]) == xx(two_q)
    assert info3.n_moves <= r and info3.n_ops_1q == n * (r + 1)
    n_swaps3 = sum(1 for k, *_ in ops3 if k == 5)
    n_swaps = sum(1 for k, *_ in two_q if k == 5)
    if d == 1:
        assert info3.n_ops_2q == n2q and n_swaps3 == 0
    else:
        assert info3.n_ops_2q < n2q and n_swaps3 <= n_swaps
        if d >= 3:
            assert n_swaps3 < n_swaps            # walk out once / back once instead of per interaction
    _replay(plan3)


@pytest.mark.parametrize("n,r,d,n2q,depth,n2q_paired,depth_paired",
                         [(10, 2, 1, 18, 4, 18, 4), (50, 2, 2, 386, 28, 290, 14), (100, 2, 2, 786, 28, 590, 14),
                          (165, 2, 1, 328, 4, 328, 4)])
def test_parallel_plan_levels(qk, n, r, d, n2q, depth, n2q_paired, depth_paired):
    """QK_PLAN_PARALLEL: no gauge moves, ops grouped into levels (op.dir = level) such that the ops of a level
    touch disjoint sites; the number of levels that contain 2-qubit ops is the dependency depth of the circuit
    (independent of the number of qubits).  With QK_PLAN_LITERAL_ORDER every site sees its ops in circuit order;
    by default interleaved distance-2 pairs share one swap (same interactions, fewer ops, half the depth)."""
    def check(plan, want_2q, want_depth):
        info = plan.info()
        assert info.n_moves == 0 and info.n_ops_2q == want_2q and info.n_ops_1q == n * (r + 1)
        ops = plan.ops()
        levels = [o[4] for o in ops]
        assert levels == sorted(levels)                      # ops are emitted level by level
        per_site = {s: [] for s in range(n)}
        two_q_levels = set()
        for lv in sorted(set(levels)):
            used = set()
            for k, s, fa, fb, l, *_ in ops:
                if l != lv:
                    continue
                sites = (s, s + 1) if 3 <= k <= 5 else (s,)
                assert not (used & set(sites)), "two ops of one level share a site"
                used |= set(sites)
                for q in sites:
                    per_site[q].append((k, s, fa, fb))
                if 3 <= k <= 5:
                    two_q_levels.add(lv)
        assert len(two_q_levels) == want_depth
        return ops, per_site

    plan, gates = _plan(qk, n, r, 0.5, d, flags=qk.QK_PLAN_PARALLEL | qk.QK_PLAN_LITERAL_ORDER)
    ops_lit, per_site = check(plan, n2q, depth)
    ref_site = {s: [] for s in range(n)}                 # the circuit's own order, per site
    for nm, q, prm in gates:
        k = qk.GATE_KIND[nm]
        fa = prm[1] if prm is not None and prm[0] in ("lin", "prod") else -1
        fb = prm[2] if prm is not None and prm[0] == "prod" else -1
        for site in q:
            ref_site[site].append((k, q[0], fa, fb))
    assert per_site == ref_site
    plan2, _ = _plan(qk, n, r, 0.5, d, flags=qk.QK_PLAN_PARALLEL)
    ops_pair, _ = check(plan2, n2q_paired, depth_paired)
    xx = lambda ops: sorted((o[2], o[3]) for o in ops if o[0] == 3)   # noqa: E731
    assert xx(ops_pair) == xx(ops_lit)                   # every interaction still there exactly once


def test_plan_from_ansatz_equals_plan_from_gates(qk):
    n, r, g, d = 20, 2, 0.7, 3
    emap = oracle.entanglement_graph(n, d)
    p1 = qk.Plan.from_ansatz(n, r, g, emap, True, 1, 1e-16, 16)
    p2, _ = _plan(qk, n, r, g, d, mode=1)
    o1, o2 = p1.ops(), p2.ops()
    assert len(o1) == len(o2)
    for a, b in zip(o1, o2):
        assert a[:5] == b[:5] and abs(a[5] - b[5]) < 1e-15


def test_plan_errors(qk):
    with pytest.raises(qk.QkError) as e:
        qk.Plan(4, [("XXPhase", (0, 2), ("const", 0.1))], 0, 1e-16, 4)     # not adjacent
    assert e.value.code == -1
    with pytest.raises(RuntimeError):
        qk.Plan(4, [("CX", (0, 1), None)], 0, 1e-16, 4)                     # unknown gate (cpu:129)
    with pytest.raises(qk.QkError) as e:
        qk.Plan(4, [("H", (0,), None)], 0, 1e-16, 1024)                     # above the limit of the stage-1 kernels
    assert e.value.code == qk.QK_ERR_LIMIT
    big = qk.Plan(4, [("H", (0,), None)], 0, 1e-16, 64)                     # above 32: the large-matrix kernel's plan
    assert big.info().threads == 256 and big.info().chi_cap == 64
    with pytest.raises(qk.QkError):
        qk.Plan(4, [("Rz", (0,), ("lin", 9, 1.0))], 0, 1e-16, 4)            # feature index out of range
    with pytest.raises(qk.QkError):
        qk.Plan(4, [("H", (0,), None)], 0, 1.5, 4)                          # truncation error out of range


def test_frag_stride(qk):
    n = 6
    D = np.array([8, 8, 16, 16, 16, 8, 8], dtype=np.int32)
    expect = sum(int(D[s]) * int(D[s + 1]) * 32 for s in range(n)) + 16
    assert qk.frag_stride(n, D) == expect
    with pytest.raises(qk.QkError):
        qk.frag_stride(n, np.array([8, 8, 12, 16, 16, 8, 8], dtype=np.int32))
    assert list(qk.pad_dims([1, 2, 8, 9, 16, 17])) == [8, 8, 8, 16, 16, 24]


def test_compute_fails_loudly_without_gpu(qk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    plan, _ = _plan(qk, 6, 1, 0.5, 1, cap=4)
    with pytest.raises(qk.QkError):
        qk.simulate(plan, oracle.synthetic_features(2, 6, 0))
    from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix
    from qkmps.engine import SingleComm
    ans = KernelStateAnsatz(6, 1, 0.5, [(0, 1)])
    with pytest.raises(qk.QkError):
        build_kernel_matrix(SingleComm(), ans, oracle.synthetic_features(2, 6, 0), truncation_error=1e-16)


def test_ansatz_mirror_matches_oracle_and_reference_errors():
    from cpu_backend.kernel_state_ansatz import KernelStateAnsatz as CpuAnsatz
    from gpu_backend.kernel_state_ansatz import KernelStateAnsatz as GpuAnsatz
    from qkmps.ansatz import structural_chi_bound
    from qkmps.synth import entanglement_graph, synthetic_features
    for n, r, d in [(10, 2, 1), (50, 2, 2), (34, 1, 3)]:
        emap = oracle.entanglement_graph(n, d)
        assert entanglement_graph(n, d) == emap
        a = GpuAnsatz(n, r, 0.5, emap)
        assert a.ansatz_circ.n_qubits == n and len(a.feature_symbol_list) == n
        ref = oracle.ansatz_gate_list(n, r, 0.5, emap)
        assert [(g[0], tuple(g[1]), g[2]) for g in a.ansatz_circ.get_commands()] == [(g[0], tuple(g[1]), g[2]) for g in ref]
    assert np.array_equal(synthetic_features(5, 7, 3), oracle.synthetic_features(5, 7, 3))
    c = CpuAnsatz(4, 1, 0.5, [(0, 1), (1, 3)])
    x = [0.1, 0.2, 0.3, 0.4]
    assert c.circuit_for_data(x) == oracle.bind_gate_list(oracle.ansatz_gate_list(4, 1, 0.5, [(0, 1), (1, 3)]), x)
    with pytest.raises(RuntimeError):
        c.circuit_for_data([0.1])
    assert structural_chi_bound(50, 2, oracle.entanglement_graph(50, 2)) == 16
    assert structural_chi_bound(20, 2, oracle.entanglement_graph(20, 1)) == 4


def test_header_is_plain_c_and_matches_the_library(qk, tmp_path):
    """include/qkmps.h is the drop-in boundary: it must compile as C99 (no C++ or torch types in the signatures), and a
    C program that references every declared function must link against libqkmps.so."""
    import re
    import subprocess
    root = pathlib.Path(__file__).resolve().parent.parent
    hdr = (root / "include" / "qkmps.h").read_text()
    names = sorted(set(re.findall(r"\b(qk_[a-z0-9_]+)\s*\(", hdr)) - {"qk_plan_create"})
    assert set(names) == set(qk.EXPORTS), (set(names) ^ set(qk.EXPORTS))
    src = tmp_path / "use_abi.c"
    src.write_text('#include "qkmps.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t f[] = {'
                   + ", ".join(f"(fn_t){n}" for n in names) +
                   '};\n  printf("%d %d\\n", (int)(sizeof(f) / sizeof(f[0])), qk_version());\n  return 0;\n}\n')
    exe = tmp_path / "use_abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{root / 'include'}", str(src), "-o", str(exe),
                           str(qk.LIB_PATH), f"-Wl,-rpath,{qk.LIB_PATH.parent}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == [str(len(names)), str(qk.lib().qk_version())]
