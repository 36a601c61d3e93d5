"""Experiment: throughput of the batched-GEMM overlap sweep (qk_big_gemm_kernel) on random MPS of a given bond dimension.
usage: python scripts/gemm_probe.py [chi] [n_sites] [n_states]"""
import pathlib, sys, time
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200"))
import qkmps
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 96
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = int(sys.argv[3]) if len(sys.argv) > 3 else 64
rng = np.random.default_rng(0)
dims = [1] + [min(chi, 2 ** min(b, n - b)) for b in range(1, n)] + [1]
states = []
for i in range(N):
    ts = []
    for s in range(n):
        a = (rng.standard_normal((dims[s], 2, dims[s + 1])) + 1j * rng.standard_normal((dims[s], 2, dims[s + 1]))) / np.sqrt(4.0 * dims[s])
        ts.append(a)
    states.append(ts)
b = qkmps.import_batch(states)
flops = 0.0
for s in range(n):
    flops += 8.0 * (2 * dims[s] * dims[s] * dims[s + 1] + 2 * dims[s] * dims[s + 1] * dims[s + 1])
flops *= N * N
for rep in range(3):
    K, ms = b.gram_store()
    print(f"chi {chi} sites {n} states {N} ({N*N} pairs): {ms:.2f} ms, {flops/ms/1e9:.2f} TFLOP/s", flush=True)
# spot check one entry against numpy
def inner(y, x):
    e = np.ones((1, 1), dtype=complex)
    for ay, ax in zip(y, x):
        t = (e @ ax.reshape(ax.shape[0], -1)).reshape(ay.shape[0] * 2, ax.shape[2])
        e = ay.reshape(ay.shape[0] * 2, ay.shape[2]).conj().T @ t
    return e[0, 0]
ref = abs(inner(states[3], states[5])) ** 2
print("entry check", K[3, 5], ref, abs(K[3, 5] - ref) / ref)
