"""CPU oracle for the quantum-kernel hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in plain numpy, the algorithm of the reference
(mmetcalf14/qml-cutensornet) for its quantum-kernel path:

  stage 1  ansatz circuit -> MPS with SVD bond truncation
  stage 2  Gram matrix K[y, x] = |<psi(Y[y]) | psi(X[x])>|^2

It is the *checker* for the CUDA path and the CPU baseline that bench.py
times.  Nothing in the product package (``qml-cutensornet_b200/``) may import
it: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do.

PARITY UNPINNED: the reference ships no tests, golden vectors or stored Gram
matrices, and all of its arithmetic lives in
third-party packages that are not vendored and not installable offline
(ITensors.jl 0.3.51 / NDTensors 0.2.21 on Julia 1.9.4 --
KernelPkg/Manifest.toml:3,288-292,461-465; pytket-cutensornet 0.6.0 --
README.md:32).  The restatement therefore follows the reference's own call
sites plus the published algorithms of those packages, and is pinned instead
by (a) the closed form for an empty entanglement map, (b) an exact
statevector simulation for <= 20 qubits, (c) the literal gate matrices of
KernelPkg/src/KernelPkg.jl:8-42, and (d) invariants (symmetry, unit diagonal,
PSD, structural bond bound).  See tests/test_oracle_*.py.
"""

from .ansatz import (  # noqa: F401
    entanglement_graph,
    ansatz_gate_list,
    bind_gate_list,
    gate_matrix,
)
from .statevector import statevector_for_data, statevector_gram  # noqa: F401
from .mps_ref import RefMPS, simulate_mps, mps_inner, truncate_itensors, truncate_pytket  # noqa: F401
from .gram_ref import compute_tile, gram_matrix  # noqa: F401
from .synth import synthetic_features  # noqa: F401
