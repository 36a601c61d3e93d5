import pathlib, sys, time
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle, qkmps
from oracle.gram_ref import simulate_batch, gram_from_mps
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz
n, r, g, d = int(sys.argv[1]), 4, float(sys.argv[2]), 4
cap = int(sys.argv[3])
X = oracle.synthetic_features(64, 165, 0)[:2, :n] if len(sys.argv) < 5 else oracle.synthetic_features(64, n, 0)[:4]
emap = oracle.entanglement_graph(n, d)
ans = KernelStateAnsatz(n, r, g, emap)
gates = ans.ansatz_circ.get_commands()
for mode, mname in ((0, "itensors"), (1, "pytket")):
    ref = simulate_batch(n, r, g, emap, X, mode=mname)
    refchi = np.array([[1] + m.bond_dims() + [1] for m in ref])
    for flags, fname in ((qkmps.QK_PLAN_LITERAL_ORDER, "literal"), (0, "default")):
        t0 = time.time()
        b = qkmps.simulate(qkmps.Plan(n, gates, mode, 1e-16, cap, flags), X)
        info = b.info()
        K, _ = b.gram_store()
        print(mname, fname, "time %.2f" % (time.time() - t0), "flags", info["flags"], "sweeps", info["sweeps"], "max chi gpu", info["chi"].max(axis=1),
              "ref", refchi.max(axis=1), "chi equal", np.array_equal(info["chi"], refchi), "K err", np.abs(K - gram_from_mps(ref)).max(), flush=True)
        if not np.array_equal(info["chi"], refchi):
            print("  gpu chi[0]", info["chi"][0].tolist()); print("  ref chi[0]", refchi[0].tolist())
