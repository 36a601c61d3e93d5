"""Small large-matrix stage-1 run for profiling: 20 qubits, 4 layers, distance 4, gamma 0.5, bond cap 128, 2 datapoints."""
import pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200"))
import qkmps
from qkmps.synth import entanglement_graph, synthetic_features
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz
n = 20
X = synthetic_features(64, 165, 0)[:2, :n]
ans = KernelStateAnsatz(n, 4, 0.5, entanglement_graph(n, 4))
plan = qkmps.Plan(n, ans.ansatz_circ.get_commands(), 0, 1e-16, 128, 0)
t0 = time.time()
b = qkmps.simulate(plan, X)
info = b.info()
print("sim_ms %.1f" % b.sim_ms(), "max chi", info["chi"].max(axis=1), "sweeps", info["sweeps"], "flags", info["flags"], "n_ops_2q", plan.info().n_ops_2q)
