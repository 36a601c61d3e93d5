"""Host-side profile of the reference-facing call (C3 workload): where the time outside the kernels goes."""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "qml-cutensornet_b200"))
import numpy as np, torch
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix
from qkmps.engine import SingleComm
from qkmps.synth import entanglement_graph, synthetic_features
n, r, g, d, N = 50, 2, 1.0, 2, 1000
X = synthetic_features(N, n, 0)
ans = KernelStateAnsatz(n, r, g, entanglement_graph(n, d))
for _ in range(3):
    build_kernel_matrix(SingleComm(), ans, X, truncation_error=1e-16)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    K = build_kernel_matrix(SingleComm(), ans, X, truncation_error=1e-16)
pr.disable()
torch.cuda.synchronize()
print("per call ms", (time.perf_counter() - t0) / 5 * 1e3, build_kernel_matrix.last_profile.get("sim_ms_x"), build_kernel_matrix.last_profile.get("gram_ms"))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
