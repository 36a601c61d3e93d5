"""Memory evolution of one circuit (the role of the reference's main_track_mem.py): prints one
``MPS size (MiB)=<value>`` line per 2-qubit operation, the format runs/mem_evol/plot.py:12-15 parses.

    python main_track_mem.py <num_features> <layers> <gamma> <distance> [circ_seed=0]
"""
import sys

import numpy as np

from gpu_backend.kernel_state_ansatz import KernelStateAnsatz
from qkmps import QK_TRUNC_PYTKET, Plan, QkError, simulate_trace
from qkmps.ansatz import structural_chi_bound
from qkmps.synth import entanglement_graph, synthetic_features


def track(num_features, reps, gamma, distance, seed=0, out=sys.stderr):
    emap = entanglement_graph(num_features, distance)
    ansatz = KernelStateAnsatz(num_features, reps, gamma, emap)
    x = synthetic_features(8, num_features, seed)[0]
    cap = int(min(16, max(1, structural_chi_bound(num_features, reps, emap))))
    while True:
        plan = Plan(num_features, ansatz.ansatz_circ.get_commands(), QK_TRUNC_PYTKET, 1e-16, cap)
        batch, trace = simulate_trace(plan, x)
        if not (batch.info()["flags"][0] & 1):
            break
        if cap >= 32:
            raise QkError(-3, "bond dimension above the shared-memory-resident limit")
        cap = min(32, (cap * 3 // 2 + 3) // 4 * 4)
    sizes = []
    for kind, site, mib in trace:
        if kind in (3, 4, 5):            # XXPhase / ZZPhase / SWAP
            print(f"MPS size (MiB)={mib}", file=out)
            sizes.append(mib)
    return np.array(sizes), batch


if __name__ == "__main__":
    a = sys.argv[1:]
    if len(a) < 4:
        raise ValueError(__doc__)
    sizes, _ = track(int(a[0]), int(a[1]), float(a[2]), int(a[3]), int(a[4]) if len(a) > 4 else 0)
    print(f"{len(sizes)} two-qubit ops, peak {sizes.max():.4f} MiB, final {sizes[-1]:.4f} MiB")
