"""Writes tests/golden/parity_ref_<workload>.npz: the oracle side of bench.py's parity check (complex overlaps among a
sample of states + their ITensors-rule bond dimensions) for workloads whose oracle simulation takes minutes
(BASELINE config 4: 165 qubits, chi ~ 100: ~3 min per circuit).  bench.py loads the file instead of re-running the
oracle when its inputs match (workload tuple + sample indices are stored and compared).

    python tests/golden/make_parity_refs.py c4 4
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200"))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n_states = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ref = bench.parity_reference(name, n_states, n_itensors=min(2, n_states))
out = pathlib.Path(__file__).resolve().parent / f"parity_ref_{name}.npz"
np.savez_compressed(out, workload=np.array(bench.WORKLOADS[name], dtype=np.float64), n_states=n_states,
                    sx=ref["sx"], sy=np.array([]) if ref["sy"] is None else ref["sy"], overlap=ref["overlap"],
                    chi_itensors=ref["chi_itensors"], sx_itensors=ref["sx_itensors"])
print(out, ref["overlap"].shape, "max chi (ITensors rule)", ref["chi_itensors"].max())
