"""Experiment: does ordering the states by stage-2 cost (sum chi^3) shorten the Gram kernel (less imbalance inside a CTA tile)?"""
import pathlib, sys
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200")); sys.path.insert(0, str(ROOT))
import torch, bench, qkmps
from qkmps.engine import SingleComm, build_gram
from qkmps.synth import entanglement_graph
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz
n, r, d, g, N, M = bench.WORKLOADS["c3"]
X, _ = bench.workload_inputs("c3")
ans = KernelStateAnsatz(n, r, g, entanglement_graph(n, d))
gates = ans.ansatz_circ.get_commands()
plans = {}
def pf(cap, early=False, parallel=False):
    k = (cap, early, parallel)
    if k not in plans:
        plans[k] = qkmps.Plan(n, gates, 1, 1e-16, cap, (2 if early else 0) | (8 if parallel else 0))
    return plans[k]
def run(Xh, tag):
    Xd = torch.from_numpy(np.ascontiguousarray(Xh)).cuda()
    for rep in range(3):
        K, prof = build_gram(SingleComm(), pf, n, Xd, None, chi_cap=16, device=0, return_device=True, structural_cap=True)
    print(tag, "sim %.2f gram %.2f" % (prof["sim_ms_x"], prof["gram_ms"]), flush=True)
    return K, prof
K0, prof = run(X, "original order ")
chi = prof["info_x"]["chi"].astype(np.float64)
cost = (chi[:, :-1] * chi[:, 1:] * chi[:, 1:]).sum(axis=1)
for name, order in (("sorted by cost  ", np.argsort(cost)), ("sorted descending", np.argsort(-cost)), ("random          ", np.random.default_rng(0).permutation(N))):
    K1, _ = run(X[order], name)
    inv = np.argsort(order)
    K1h = K1.cpu().numpy()
    print("   max |K diff| after un-permuting:", float(np.abs(K1h[np.ix_(inv, inv)] - K0.cpu().numpy()).max()))
print("cost min/median/max", cost.min(), np.median(cost), cost.max())
