import pathlib, sys, time, os
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "qml-cutensornet_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import oracle, qkmps
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz
for (n, r, g, d, nx, cap) in [(165, 4, 0.1, 4, 64, 24), (50, 2, 1.0, 2, 64, 16), (50, 4, 0.1, 4, 64, 24), (50, 2, 0.1, 4, 64, 24), (50, 2, 0.1, 2, 64, 24), (50, 2, 1.0, 2, 64, 24)]:
    emap = oracle.entanglement_graph(n, d)
    X = oracle.synthetic_features(nx, n, 0)
    ans = KernelStateAnsatz(n, r, g, emap)
    plan = qkmps.Plan(n, ans.ansatz_circ.get_commands(), 1, 1e-16, cap)
    b = qkmps.simulate(plan, X); b = qkmps.simulate(plan, X)
    info = b.info(); pi = plan.info()
    print(f"n={n} r={r} d={d} g={g} cap={cap} G={pi.threads}: sim {b.sim_ms():.1f} ms for {nx} dp; ops {pi.n_ops} 2q {pi.n_ops_2q}; us/op {b.sim_ms()*1e3/pi.n_ops:.1f}; sweeps/2q {info['sweeps'].mean()/pi.n_ops_2q:.2f}; chi max {info['chi'].max()} mean {info['chi'].mean():.1f}; flags {np.bitwise_or.reduce(info['flags'])}", flush=True)
