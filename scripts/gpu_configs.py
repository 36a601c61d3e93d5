"""Run BASELINE configs 4 (gamma 0.1) and 5 on one GPU; spot-check a sample against the oracle."""
import pathlib, sys, time
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "qml-cutensornet_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import oracle
from oracle.gram_ref import gram_from_mps, simulate_batch
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix
from qkmps.engine import SingleComm

def run(name, n, r, g, d, nx, ny, sample=6):
    emap = oracle.entanglement_graph(n, d)
    X = oracle.synthetic_features(nx, n, 0)
    Y = oracle.synthetic_features(ny, n, 1) if ny else None
    ans = KernelStateAnsatz(n, r, g, emap)
    build_kernel_matrix(SingleComm(), ans, X[:8], truncation_error=1e-16)  # warm-up
    t0 = time.perf_counter()
    K = build_kernel_matrix(SingleComm(), ans, X, Y, truncation_error=1e-16)
    dt = time.perf_counter() - t0
    prof = build_kernel_matrix.last_profile
    chi = prof["info_x"]["chi"]
    xs = simulate_batch(n, r, g, emap, X[:sample], mode="pytket")
    if Y is None:
        Kref = gram_from_mps(xs); Ksub = K[:sample, :sample]
    else:
        ys = simulate_batch(n, r, g, emap, Y[:sample], mode="pytket")
        Kref = gram_from_mps(xs, ys); Ksub = K[:sample, :sample]
    err = np.abs(Ksub - Kref).max()
    pi = prof["plan"]
    print(f"{name}: n={n} r={r} d={d} gamma={g} K{K.shape} wall {dt*1e3:.1f} ms  sim {prof['sim_ms_x']:.1f}+{prof['sim_ms_y']:.1f} ms "
          f"gram {prof['gram_ms']:.1f} ms  chi max {chi.max()} mean-max {chi.max(axis=1).mean():.2f} cap {prof['chi_cap']} "
          f"ops {pi.n_ops} (2q {pi.n_ops_2q}, moves {pi.n_moves})  max|K-Koracle| on {sample}x{sample} sample = {err:.2e}  "
          f"offdiag range [{np.min(K):.2e}, {np.max(K - np.eye(*K.shape) if Y is None else K):.2e}]", flush=True)
    assert err < 1e-8

run("C5", 100, 2, 1.0, 2, 1000, 1000)
run("C5 g0.1", 100, 2, 0.1, 2, 1000, 1000)
run("C4 g0.1", 165, 4, 0.1, 4, 256, 0, sample=3)
run("C2", 20, 2, 0.5, 1, 200, 0)
run("runtime_scaling shape", 165, 2, 0.1, 1, 1280, 0, sample=4)
