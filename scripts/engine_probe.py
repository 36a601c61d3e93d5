"""Experiment: host-side timeline of one streamlined build_gram step (C3 workload)."""
import pathlib, sys, time
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
Xd = torch.from_numpy(X).cuda()
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K, prof = build_gram(SingleComm(), pf, n, Xd, None, chi_cap=16, device=0, return_device=True, structural_cap=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("rep", rep, "wall %.1f ms" % ((t1 - t0) * 1e3), "sim %.1f gram %.1f" % (prof["sim_ms_x"], prof["gram_ms"]))
    print("   ", [(a, round(b, 2)) for a, b in prof["host_trace_ms"]])
