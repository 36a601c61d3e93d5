"""Experiment: stage timings of the high-bond-dimension path (BASELINE config 4 shape) on one GPU.
usage: python scripts/c4_probe.py [n_points] [gamma] [n_qubits] [chi]"""
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200"))
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix  # noqa: E402
from qkmps.engine import SingleComm  # noqa: E402
from qkmps.synth import entanglement_graph, synthetic_features  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
n = int(sys.argv[3]) if len(sys.argv) > 3 else 165
chi = int(sys.argv[4]) if len(sys.argv) > 4 else None
X = synthetic_features(max(N, 64), n, 0)[:N]
ans = KernelStateAnsatz(n, 4, g, entanglement_graph(n, 4))
for rep in range(2):
    t0 = time.perf_counter()
    K = build_kernel_matrix(SingleComm(), ans, X, truncation_error=1e-16, chi=chi)
    wall = time.perf_counter() - t0
    p = build_kernel_matrix.last_profile
    chi_max = p["info_x"]["chi"].max(axis=1)
    print(f"rep {rep}: N={N} n={n} gamma={g}: wall {wall:.2f} s, sim {p['sim_ms_x']:.0f} ms, gram {p['gram_ms']:.0f} ms, "
          f"kernel {p['gram_kernel']}, cap {p['chi_cap']}, max chi {chi_max.max()} (mean {chi_max.mean():.1f}), "
          f"sweeps/state {p['info_x']['sweeps'].mean():.0f}, diag err {np.abs(np.diag(K) - 1).max():.1e}", flush=True)
